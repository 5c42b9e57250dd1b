mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -n 4 -k "owned_tiles or partition" ) > gpurun_out/pytest_owned.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_owned.log
tail -n 3 gpurun_out/pytest_owned.log
N=${1:-2}
for h in peer; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --handoff $h --no-count > gpurun_out/bench_n${N}_$h.log 2>&1; echo "$h rc=$?"
tail -n 1 gpurun_out/bench_n${N}_$h.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms %.3f e2e %.0f (%.3f ms) path=%s host_same=%s dev_same=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['e2e_path'], d['config']['multi_gpu_host_frame_equals_single_gpu_frame'], d['config']['multi_gpu_frame_equals_single_gpu_frame']))"
grep -i "warn\|error\|leak" gpurun_out/bench_n${N}_$h.log | head -5
done
