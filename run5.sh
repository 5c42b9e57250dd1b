mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
for k in texgather texpair texpair2; do
timeout 300 python bench.py --steps 20 --warmup 3 --kernel $k --no-cpu-baseline > gpurun_out/bench_$k.json 2> gpurun_out/bench_$k.err; echo "$k rc=$?"
cut -c1-200 gpurun_out/bench_$k.json; grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_$k.json
done
VR_TEXPAIR2_MINB=4 timeout 300 python bench.py --steps 20 --warmup 3 --kernel texpair2 --no-cpu-baseline > gpurun_out/bench_texpair2_minb4.json 2> gpurun_out/bench_texpair2_minb4.err
grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_texpair2_minb4.json
for cam in K0 K1; do for k in texgather texpair texpair2; do
timeout 300 python bench.py --steps 10 --warmup 3 --kernel $k --camera $cam --no-cpu-baseline --no-count > gpurun_out/bench_${k}_$cam.json 2> gpurun_out/bench_${k}_$cam.err
echo "$cam $k $(grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_${k}_$cam.json)"
done; done
