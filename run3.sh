mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_texgather -s 3 -c 1 -o gpurun_out/r01_prof_texgather python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_packed -s 3 -c 1 -o gpurun_out/r01_prof_packed python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count --kernel fast > gpurun_out/ncu_full_p.log 2>&1
for k in fast windowed direct; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel $k > gpurun_out/bench_$k.log 2>&1; done
for cfg in "--camera K0" "--camera K1" "--alpha 1.0" "--filter nearest" "--config C3" "--config C2" "--config C1"; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg > "gpurun_out/bench_$(echo $cfg | tr -d ' -').log" 2>&1; done
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
tail -n 1 gpurun_out/ncu_full.log
