# round 2, lab 7: ingest kernel iteration: cell/stat tests, ingest timing, launch list (durations + DRAM bytes) of the upload kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -n 4 -k "cell_table or volume_stats or launch_order or partition or synthetic" ) > gpurun_out/pytest_gpu7.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu7.log
tail -n 3 gpurun_out/pytest_gpu7.log
( timeout 600 python tools/lab/ingest.py 1 ) > gpurun_out/lab_ingest_1.log 2>&1; tail -n 6 gpurun_out/lab_ingest_1.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6 --csv --log-file gpurun_out/launches_ingest.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-count --no-dense > gpurun_out/launches_ingest.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum -k regex:cta_order --clock-control none -c 3 --csv --log-file gpurun_out/launches_order.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-count --no-dense > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/launches_ingest.csv", "gpurun_out/launches_order.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
    agg = {}
    for r in rows[1:]:
        agg.setdefault((r[ii], r[ki][:60]), {})[r[mi]] = r[vi]
    for (i, k), m in agg.items():
        print(i, k, m)
PY
