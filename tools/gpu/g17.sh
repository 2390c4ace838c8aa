set -x; mkdir -p gpurun_out
timeout 600 python tools/tile_check.py --levels 0,1 --B 1024 --only dw --skip-old --iters 5 > gpurun_out/g17_tile_check.log 2>&1; echo "rc=$?" >> gpurun_out/g17_tile_check.log; cat gpurun_out/g17_tile_check.log
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "tile or bwd_w" > gpurun_out/g17_tests.log 2>&1; tail -3 gpurun_out/g17_tests.log
