mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for k in auto direct; do
timeout 600 python bench.py --steps 10 --warmup 3 --config C3 --tf --kernel $k --cpu-row-stride 8 --no-count > gpurun_out/bench_C3_tf_$k.json 2> gpurun_out/bench_C3_tf_$k.err; echo "rc=$?"; tail -2 gpurun_out/bench_C3_tf_$k.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_C3_tf_$k.json"))
print("C3 tf $k kernel %.3f ms (%s) value %.1f e2e %.1f parity %s" % (d["roofline"]["kernel_ms_avg"], d["roofline"]["kernel"], d["value"], d["e2e"]["value"], d["cpu_baseline"]["parity_bit_exact_on_sample"]))
PY
done
