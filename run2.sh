set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --impl reference > gpurun_out/bench_ref.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_ref.log
tail -n 4 gpurun_out/smoke.log gpurun_out/pytest_gpu.log
