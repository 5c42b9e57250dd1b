# round 2, GPU call 2: parity of the leaping skip kernels + lab 2 (depth / occupancy of nearest + skipping forms)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -n 4 ) > gpurun_out/pytest_parity2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity2.log
tail -n 12 gpurun_out/pytest_parity2.log
( timeout 1200 python tools/lab/variants.py ) > gpurun_out/lab_variants2.log 2>&1; echo "rc=$?" >> gpurun_out/lab_variants2.log
tail -n 3 gpurun_out/lab_variants2.log
