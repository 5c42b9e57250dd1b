mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("default: value %.1f kernel %.3f e2e %.1f (%.3f ms) traffic %s parity %s clocks %s launches %s" % (d["value"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["traffic"], d["cpu_baseline"]["parity_bit_exact_on_sample"], d["clocks"], d["gpu_launches"]))
PY
