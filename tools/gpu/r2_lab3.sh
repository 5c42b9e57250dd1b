# round 2, lab 3: CTA shape / resident warps; parity first (includes the strided band copy)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -n 4 ) > gpurun_out/pytest_parity3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity3.log
tail -n 4 gpurun_out/pytest_parity3.log
( timeout 1200 python tools/lab/variants.py ) > gpurun_out/lab_variants3.log 2>&1; echo "rc=$?" >> gpurun_out/lab_variants3.log
tail -n 3 gpurun_out/lab_variants3.log
