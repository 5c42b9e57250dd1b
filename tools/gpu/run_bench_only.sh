mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("default: value %.1f kernel %.3f e2e %.1f (%.3f ms) traffic %s parity %s clocks %s launches %s" % (d["value"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["traffic"], d["cpu_baseline"]["parity_bit_exact_on_sample"], d["clocks"], d["gpu_launches"]))
PY
timeout 300 python bench.py --filter nearest --no-cpu-baseline --no-count > gpurun_out/bench_nearest_auto.json 2> gpurun_out/bench_nearest_auto.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_nearest_auto.json"))
print("nearest: value %.1f kernel %.3f (%s) e2e %.1f (%.3f ms)" % (d["value"], d["roofline"]["kernel_ms_avg"], d["roofline"]["kernel"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
