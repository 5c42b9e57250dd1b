# round 2, lab 5: GPU-built launch-order table: full GPU test suite, LPT lab (off vs product), default bench line
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu5.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu5.log
tail -n 5 gpurun_out/pytest_gpu5.log
rm -f gpurun_out/lab_lpt.log
for m in 0 auto; do ( timeout 600 python tools/lab/lpt.py $m ) >> gpurun_out/lab_lpt.log 2>&1; done
cat gpurun_out/lab_lpt.log
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -c 3000 gpurun_out/bench5.json
