# round 2, final single-GPU sanity: full GPU suite + stream-aliasing lab
mkdir -p gpurun_out
( timeout 100 python -m pytest tests -m gpu -x -q -n 6 ) > gpurun_out/pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -n 4 gpurun_out/pytest_gpu_final.log
( timeout 25 python tools/lab/streams.py 0; timeout 25 python tools/lab/streams.py 6; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 25 python tools/lab/streams.py 6 ) > gpurun_out/lab_streams.log 2>&1
cat gpurun_out/lab_streams.log
