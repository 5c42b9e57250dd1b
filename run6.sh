mkdir -p gpurun_out
for k in texpair texpair2; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_$k -s 3 -c 1 -o gpurun_out/ncu_$k -f python bench.py --steps 2 --warmup 3 --kernel $k --no-cpu-baseline --no-count > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
