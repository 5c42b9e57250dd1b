#!/usr/bin/env python
"""Condenses an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers
DESIGN.md / bench.py quote.  Usage: python profiles/summarize_ncu.py gpurun_out/x.ncu-rep [out.txt]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    # the texture path: what the whole design argument rests on
    "l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__tex_writeback_active.sum",
    "l1tex__data_pipe_tex_wavefronts.sum", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_tex.sum", "l1tex__t_sectors_pipe_tex.sum", "l1tex__t_sector_pipe_tex_hit_rate.pct",
    "smsp__inst_executed_pipe_tex.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_pred_on.sum",
    "smsp__warps_active.avg.per_cycle_active", "sm__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
    "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                out.append(f"{k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[h]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        out.append("warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    # optional: record the DRAM traffic of the (last) kernel under a workload key for bench.py's roofline.traffic
    if len(sys.argv) > 4:
        import json
        import os
        tj, key = sys.argv[3], sys.argv[4]
        d = dict(zip(hdr, rows[-1]))

        def num(k):
            try:
                v = float(d[k].replace(",", ""))
            except Exception:
                return 0.0
            u = units[hdr.index(k)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        ent = {"kernel": d.get("Kernel Name", "?").split("(")[0].split("<")[0].replace("void vr::", "").replace("void ", "").replace("vr::", "").strip(),
               "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
               "source": os.path.basename(rep)}
        allj = json.load(open(tj)) if os.path.exists(tj) else {}
        allj[key] = ent
        json.dump(allj, open(tj, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
