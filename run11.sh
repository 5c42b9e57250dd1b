mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_peer_frame.py -m gpu -x -q ) > gpurun_out/pytest_peer.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_peer.log
tail -n 6 gpurun_out/pytest_peer.log
for h in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --handoff $h --no-count > gpurun_out/bench_n2_$h.log 2>&1; echo "$h rc=$?"
tail -n 1 gpurun_out/bench_n2_$h.log | cut -c1-200
done
