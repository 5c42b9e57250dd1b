# round 2, lab 10: histogram with a short LUT: stats tests + launch list of the upload kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_renderer_core.py -m gpu -x -q -n 4 -k "cell_table or volume_stats or stats or histogram or synthetic or core" ) > gpurun_out/pytest_gpu10.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu10.log
tail -n 3 gpurun_out/pytest_gpu10.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 5 --csv --log-file gpurun_out/launches_ingest.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-count --no-dense > gpurun_out/launches_ingest.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_ingest.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
agg = {}
for r in rows[1:]:
    agg.setdefault((r[ii], r[ki][:60]), {})[r[mi]] = r[vi]
for (i, k), m in agg.items():
    print(i, k, m)
PY
