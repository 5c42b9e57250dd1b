mkdir -p gpurun_out
for v in 0 1 2; do
if [ $v = 0 ]; then unset VOLREN_B200_LIB; else export VOLREN_B200_LIB=$PWD/volume-renderer_b200/lib/libvolren_b200_imm$v.so; fi
for cam in K2 K0; do
timeout 300 python bench.py --steps 20 --warmup 3 --camera $cam --no-count --cpu-row-stride 8 > gpurun_out/bench_imm${v}_$cam.json 2> gpurun_out/bench_imm${v}_$cam.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_imm${v}_$cam.json"))
print("imm$v $cam kernel %.3f ms value %.1f parity %s" % (d["roofline"]["kernel_ms_avg"], d["value"], d["cpu_baseline"]["parity_bit_exact_on_sample"]))
PY
done; done
