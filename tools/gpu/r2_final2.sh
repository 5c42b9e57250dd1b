# round 2, last call: full GPU suite on the final code
mkdir -p gpurun_out
( timeout 70 python -m pytest tests -m gpu -x -q -n 6 ) > gpurun_out/pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -n 4 gpurun_out/pytest_gpu_final.log
