mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu2.log
tail -n 5 gpurun_out/pytest_gpu2.log
for mb in 6 5; do
VR_PIPE_MINB=$mb timeout 300 python bench.py --steps 20 --warmup 3 --kernel texpair_pipe --no-cpu-baseline --no-count > gpurun_out/bench_pipe_$mb.json 2> gpurun_out/bench_pipe_$mb.err; echo "pipe minb=$mb rc=$? $(grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_pipe_$mb.json)"
done
for cam in K0 K1; do
timeout 300 python bench.py --steps 10 --warmup 3 --kernel texpair_pipe --camera $cam --no-cpu-baseline --no-count > gpurun_out/bench_pipe_$cam.json 2> gpurun_out/bench_pipe_$cam.err
echo "$cam pipe $(grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_pipe_$cam.json)"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_texpair_pipe -s 3 -c 1 -o gpurun_out/ncu_texpair_pipe -f python bench.py --steps 2 --warmup 3 --kernel texpair_pipe --no-cpu-baseline --no-count > gpurun_out/ncu_pipe.log 2>&1; echo "ncu rc=$?"
