mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu3.log
tail -n 5 gpurun_out/pytest_gpu3.log
for k in texpair_pipe hybrid zlsu; do for cam in K2 K0 K1; do
timeout 300 python bench.py --steps 20 --warmup 3 --kernel $k --camera $cam --no-cpu-baseline --no-count > gpurun_out/bench_${k}_$cam.json 2> gpurun_out/bench_${k}_$cam.err; echo "$k $cam rc=$? $(grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_${k}_$cam.json)"
done; done
for k in hybrid zlsu; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_texpair_pipe -s 3 -c 1 -o gpurun_out/ncu_$k -f python bench.py --steps 2 --warmup 3 --kernel $k --no-cpu-baseline --no-count > gpurun_out/ncu_$k.log 2>&1; echo "ncu rc=$?"
done
