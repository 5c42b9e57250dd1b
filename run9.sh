mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2.txt 2>&1
for h in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --handoff $h > gpurun_out/bench_n2_$h.log 2>&1; echo "$h rc=$?"
tail -n 1 gpurun_out/bench_n2_$h.log | cut -c1-900
done
