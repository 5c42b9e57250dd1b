# round 2, lab 4: (a) CTA launch order on small grids (LPT), (b) 70-register variants of the trilinear kernel
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -n 4 -x ) > gpurun_out/pytest_parity4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity4.log
tail -n 3 gpurun_out/pytest_parity4.log
for m in 0 1 auto; do ( timeout 600 python tools/lab/lpt.py $m ) >> gpurun_out/lab_lpt.log 2>&1; done
cat gpurun_out/lab_lpt.log
( LAB_COMBOS="4,32,42;4,24,42;3,24,42" timeout 1200 python tools/lab/variants.py ) > gpurun_out/lab_variants4.log 2>&1; echo "rc=$?" >> gpurun_out/lab_variants4.log
cat gpurun_out/lab_variants4.log
