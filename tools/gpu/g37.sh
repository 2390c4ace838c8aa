set -x; mkdir -p gpurun_out
timeout 300 python tools/tile_check.py --levels 0 --B 1024 --iters 5 --only out,bwo > gpurun_out/g37_tile.log 2>&1; cat gpurun_out/g37_tile.log
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/g37_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g37_tests.log; tail -3 gpurun_out/g37_tests.log
