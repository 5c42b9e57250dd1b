mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for k in auto direct; do
timeout 600 python bench.py --steps 10 --warmup 3 --mip --alpha 0.5 --kernel $k --cpu-row-stride 8 --no-count > gpurun_out/bench_mip_$k.json 2> gpurun_out/bench_mip_$k.err; echo "rc=$?"; tail -2 gpurun_out/bench_mip_$k.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_mip_$k.json"))
print("C4 mip $k kernel %.3f ms (%s) value %.1f e2e %.1f (%.3f ms) parity %s" % (d["roofline"]["kernel_ms_avg"], d["roofline"]["kernel"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["parity_bit_exact_on_sample"]))
PY
done
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("default: value %.1f kernel %.3f e2e %.1f (%.3f ms) traffic %s parity %s clocks %s" % (d["value"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["traffic"], d["cpu_baseline"]["parity_bit_exact_on_sample"], d["clocks"]))
PY
