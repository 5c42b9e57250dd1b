# round 2, GPU call 1: parity of the restructured kernels (quick subset) + the pipeline-depth / skipping lab
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -n 4 -x ) > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity.log
tail -n 15 gpurun_out/pytest_parity.log
( timeout 1200 python tools/lab/variants.py ) > gpurun_out/lab_variants.log 2>&1; echo "rc=$?" >> gpurun_out/lab_variants.log
tail -n 5 gpurun_out/lab_variants.log
