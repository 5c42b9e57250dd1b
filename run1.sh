set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 300 python bench.py --steps 5 --warmup 3 --camera K0 --no-cpu-baseline > gpurun_out/bench_k0.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --alpha 1.0 --no-cpu-baseline > gpurun_out/bench_a1.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --filter nearest --no-cpu-baseline > gpurun_out/bench_nearest.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_direct -s 1 -c 1 -o gpurun_out/prof_direct python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-count > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench.log
