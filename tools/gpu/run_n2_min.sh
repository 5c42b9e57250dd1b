mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-count > gpurun_out/bench_n2_peer.log 2>&1; echo "rc=$?"
tail -n 1 gpurun_out/bench_n2_peer.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms %.3f e2e %.0f (%.3f ms) host_same=%s dev_same=%s clocks=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['multi_gpu_host_frame_equals_single_gpu_frame'], d['config']['multi_gpu_frame_equals_single_gpu_frame'], d['clocks']))"
