mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_renderer_core.py -m gpu -x -q -n 4 ) > gpurun_out/pytest_ingest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ingest.log
tail -n 4 gpurun_out/pytest_ingest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ingest.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count > gpurun_out/launches_ingest.log 2>&1; echo "ncu rc=$?"
grep -E "minmax|histogram|pad_volume|zpair_pack" gpurun_out/launches_ingest.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | sort | uniq -c | head
