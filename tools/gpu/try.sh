set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mse or lap" > gpurun_out/try_tests.log 2>&1; echo "rc=$?" >> gpurun_out/try_tests.log; tail -5 gpurun_out/try_tests.log
timeout 300 python tools/microbench.py --batches 1024 --only none --iters 3 > /dev/null 2>&1
