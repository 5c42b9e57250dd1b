mkdir -p gpurun_out
for cfg in C1 C2 C3 C5; do for k in texgather texpair auto; do
timeout 600 python bench.py --steps 10 --warmup 3 --config $cfg --kernel $k --no-cpu-baseline --no-count > gpurun_out/bench_${cfg}_$k.json 2> gpurun_out/bench_${cfg}_$k.err
echo "$cfg $k rc=$? $(grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_${cfg}_$k.json) $(grep -o '"kernel": "[a-z_0-9]*"' gpurun_out/bench_${cfg}_$k.json | head -1)"
done; done
for a in 1.0; do for k in texgather auto; do
timeout 600 python bench.py --steps 10 --warmup 3 --alpha $a --kernel $k --no-cpu-baseline --no-count > gpurun_out/bench_alpha${a}_$k.json 2> gpurun_out/bench_alpha${a}_$k.err
echo "alpha $a $k $(grep -o '"kernel_ms_avg": [0-9.]*' gpurun_out/bench_alpha${a}_$k.json)"
done; done
tail -3 gpurun_out/bench_C5_auto.err
