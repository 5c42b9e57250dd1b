# round 2 multi-GPU, peer hand-off only (two target frames, asynchronous marches): N GPUs; N = 8 adds C5 and the windowed frame
N=$1
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { name=$1; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --handoff peer "$@" > $O/bench_n${N}_$name.log 2>&1; echo "$name rc=$?"
tail -n 1 $O/bench_n${N}_$name.log > $O/bench_n${N}_$name.json
python - <<PY
import json
d=json.loads(open("$O/bench_n${N}_$name.json").read())
print("N=$N $name: value %.1f Mrays/s (%.3f ms/step) kernel %.3f ms e2e %.1f (%.3f ms) same=%s hostsame=%s launches %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["multi_gpu_frame_equals_single_gpu_frame"], d["config"]["multi_gpu_host_frame_equals_single_gpu_frame"], d["gpu_launches"]))
PY
}
run peer
if [ "$N" = "8" ]; then
run C5 --config C5
run window --window 1000 3000 --alpha 0.05
fi
