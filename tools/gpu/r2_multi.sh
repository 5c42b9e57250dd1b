# round 2 multi-GPU: bench.py under torchrun at N GPUs, both hand-offs (+ C5 and the peer-frame tests at N=8 / N=2)
N=$1
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi topo -m > $O/topo_$N.txt 2>&1
for h in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --handoff $h > $O/bench_n${N}_$h.log 2>&1; echo "$h rc=$?"
tail -n 1 $O/bench_n${N}_$h.log > $O/bench_n${N}_$h.json
python - <<PY
import json
d=json.loads(open("$O/bench_n${N}_$h.json").read())
print("N=$N $h: value %.1f Mrays/s (%.3f ms/step) kernel %.3f ms e2e %.1f (%.3f ms) same=%s hostsame=%s launches %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["multi_gpu_frame_equals_single_gpu_frame"], d["config"]["multi_gpu_host_frame_equals_single_gpu_frame"], d["gpu_launches"]))
PY
done
if [ "$N" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --config C5 > $O/bench_n${N}_C5.log 2>&1; echo "C5 rc=$?"
tail -n 1 $O/bench_n${N}_C5.log > $O/bench_n${N}_C5.json; cut -c1-200 $O/bench_n${N}_C5.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --window 1000 3000 --alpha 0.05 > $O/bench_n${N}_window.log 2>&1; echo "window rc=$?"
tail -n 1 $O/bench_n${N}_window.log > $O/bench_n${N}_window.json; cut -c1-200 $O/bench_n${N}_window.json
fi
if [ "$N" = "2" ]; then
( timeout 600 python -m pytest tests/test_gpu_peer_frame.py -m gpu -q ) > $O/pytest_peer_n2.log 2>&1; echo "peer tests rc=$?"; tail -n 3 $O/pytest_peer_n2.log
fi
