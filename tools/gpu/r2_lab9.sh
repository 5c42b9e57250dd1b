# round 2, lab 9: one ncu --set full capture of the fused ingest kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pad_cells -c 1 -f -o gpurun_out/ncu_pad_cells python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-count --no-dense > gpurun_out/ncu_pad_cells.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_ncu.py gpurun_out/ncu_pad_cells.ncu-rep gpurun_out/ncu_pad_cells_summary.txt > /dev/null 2>&1
cat gpurun_out/ncu_pad_cells_summary.txt
ncu -i gpurun_out/ncu_pad_cells.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=['smsp__average_warp_latency_issue_stalled','smsp__average_warps_issue_stalled','l1tex__data_pipe_lsu_wavefronts_mem_shared','smsp__inst_executed_op_shared_atom','sm__warps_active.avg.pct','dram__throughput','lts__t_sectors_op_write','lts__t_sectors_op_read','lts__t_sectors_op_atom','lts__t_sectors_op_red','smsp__warp_issue_stalled']
for a,b in zip(h,v):
    if any(w in a for w in want): print(a,b)
" | head -60
