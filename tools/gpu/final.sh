set -x; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gpu_tests.log; tail -3 gpurun_out/r02_gpu_tests.log
bash tools/gpu/measure_1gpu.sh
