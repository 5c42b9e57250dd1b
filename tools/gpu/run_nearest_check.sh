mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for k in auto fast; do
timeout 300 python bench.py --steps 20 --warmup 3 --filter nearest --kernel $k --cpu-row-stride 16 --no-count > gpurun_out/bench_nearest_$k.json 2> gpurun_out/bench_nearest_$k.err; echo "rc=$?"; tail -1 gpurun_out/bench_nearest_$k.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_nearest_$k.json"))
print("nearest $k kernel %.3f ms (%s) value %.1f e2e %.1f parity %s" % (d["roofline"]["kernel_ms_avg"], d["roofline"]["kernel"], d["value"], d["e2e"]["value"], d["cpu_baseline"]["parity_bit_exact_on_sample"]))
PY
done
