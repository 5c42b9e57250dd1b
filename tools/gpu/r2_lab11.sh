# round 2, lab 11: pipelined host path (vr_render_submit / vr_render_wait): tests + default bench line
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -n 4 ) > gpurun_out/pytest_gpu11.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu11.log
tail -n 15 gpurun_out/pytest_gpu11.log
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -n 3 gpurun_out/bench11.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench11.json").read().strip().splitlines()[-1])
print("value %.1f (%.3f ms) e2e %s host_frames_equal %s" % (d["value"], d["ms_per_step"], d["e2e"], d["config"].get("host_frames_equal_device_frame")))
PY
