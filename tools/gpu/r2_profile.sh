# round 2: ncu evidence.  Launch list of the default bench + one --set full capture per workload/kernel of interest;
# summaries and traffic.json go to gpurun_out/r02 (copied to profiles/r02/).
mkdir -p gpurun_out/r02
O=gpurun_out/r02
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_ingest.csv python -c "
import sys; sys.path[:0]=['volume-renderer_b200/python']
import volren_b200 as vb
from volren_b200 import workloads
with vb.Context(64, 64) as ctx:
    ctx.upload_synthetic((1024, 1024, 1024), 2, 4095, workloads.SEEDS['C4'])
    ctx.set_camera(workloads.camera_block('K2'))
    ctx.set_params(vb.default_params(alpha_scale=0.02, min_val=1000, max_val=3000, filter=1)); ctx.render()
    ctx.set_camera(workloads.camera_block('K1')); ctx.render()
    ctx.set_params(vb.default_params(alpha_scale=0.02, min_val=1000, max_val=3000, filter=0)); ctx.render()
" > $O/launches_ingest.log 2>&1; echo "ingest launch list rc=$?"
cap() {  # name kernel-regex workload-key bench-args...
  name=$1; kre=$2; key=$3; shift 3
  timeout 900 $NCU -k regex:$kre -s 3 -c 1 -f -o $O/ncu_$name python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count --no-dense "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  python profiles/summarize_ncu.py $O/ncu_$name.ncu-rep $O/ncu_${name}_summary.txt $O/traffic.json "$key" > /dev/null 2>&1
  head -n 3 $O/ncu_${name}_summary.txt
  # gpurun_out/ is limited to 64 MiB: keep only the headline capture (source page), drop the other reports
  if [ "$name" != "texpair_C4_K2" ]; then rm -f $O/ncu_$name.ncu-rep; else ncu -i $O/ncu_$name.ncu-rep --page source --csv > $O/ncu_${name}_source.csv 2>/dev/null; fi
}
cap texpair_C4_K2      march_texpair_kernel "C4/K2/trilinear/0.02/0-4095"
cap texpair_C4_K0      march_texpair_kernel "C4/K0/trilinear/0.02/0-4095" --camera K0
cap texpair_C4_alpha1  march_texpair_kernel "C4/K2/trilinear/1.0/0-4095" --alpha 1.0
cap nearest_C4_K2      march_nearest_kernel "C4/K2/nearest/0.02/0-4095" --filter nearest
cap texpair_C4_window_skip march_texpair_kernel "C4/K2/trilinear/0.05/1000-3000/skip" --window 1000 3000 --alpha 0.05
cap texpair_C3_skip    march_texpair_kernel "C3/K2/trilinear/0.05/1000-3000/skip" --config C3 --alpha 0.05
cap texpair_C2         march_texpair_kernel "C2/K2/trilinear/0.02/0-255" --config C2
cap texpair_C5         march_texpair_kernel "C5/K2/trilinear/0.02/0-4095" --config C5
# the fused ingest kernel (HBM-bound: padded copy + cell table + min/max in one read of the source) and the histogram
for k in pad_cells histogram_kernel; do
  timeout 600 $NCU -k regex:$k -c 1 -f -o $O/ncu_ingest_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-count --no-dense > $O/ncu_ingest_$k.log 2>&1; echo "ncu ingest $k rc=$?"
  python profiles/summarize_ncu.py $O/ncu_ingest_$k.ncu-rep $O/ncu_ingest_${k}_summary.txt > /dev/null 2>&1; head -n 3 $O/ncu_ingest_${k}_summary.txt
  rm -f $O/ncu_ingest_$k.ncu-rep
done
ls -la $O | head -40
cuobjdump -sass -fun $(cuobjdump -elf volume-renderer_b200/lib/libvolren_b200.so | grep -o "_ZN2vr20march_texpair_kernelItLi0ELi1ELb1ELb1ELi0ELi4ELb0ELi32ELi4ELi2EEEvNS_11FrameConstsENS_9MarchArgsE" | head -1) volume-renderer_b200/lib/libvolren_b200.so > $O/sass_texpair_headline.txt 2>&1
