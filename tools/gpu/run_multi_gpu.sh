N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
for h in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --handoff $h > gpurun_out/bench_n${N}_$h.log 2>&1; echo "$h rc=$?"
tail -n 1 gpurun_out/bench_n${N}_$h.log | cut -c1-160
done
if [ "$N" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --config C5 > gpurun_out/bench_n${N}_C5.log 2>&1; echo "C5 rc=$?"
tail -n 1 gpurun_out/bench_n${N}_C5.log | cut -c1-160
fi
