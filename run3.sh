mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_fast -s 3 -c 1 -o gpurun_out/r01_prof_fast python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_windowed -s 3 -c 1 -o gpurun_out/r01_prof_windowed python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count --kernel windowed > gpurun_out/ncu_full_w.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel windowed > gpurun_out/bench_windowed.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel direct > gpurun_out/bench_direct.log 2>&1
tail -n 2 gpurun_out/ncu_full.log gpurun_out/ncu_full_w.log
