# round 2, lab 13 (2 GPUs): two march streams in the peer hand-off: peer tests (subset) + bench at N=2
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_peer_frame.py -m gpu -x -q -k "two_target or barrier" ) > gpurun_out/pytest_peer13.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_peer13.log
tail -n 4 gpurun_out/pytest_peer13.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --handoff peer > gpurun_out/bench13_n2.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench13_n2.log > gpurun_out/bench13_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench13_n2.json").read())
print("N=2 peer: value %.1f (%.3f ms/step) kernel %.3f ms e2e %.1f (%.3f ms) same=%s hostsame=%s launches %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["multi_gpu_frame_equals_single_gpu_frame"], d["config"]["multi_gpu_host_frame_equals_single_gpu_frame"], d["gpu_launches"]))
PY
