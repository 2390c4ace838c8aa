set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_tc.py -m gpu -x -q -s -k "dropin or measured or benchmarked or generate_for_opt or l1_loss or tile_staged" > gpurun_out/g12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g12_tests.log
grep -E "engine vs fp64|passed|failed|rc=|Error|error|assert" gpurun_out/g12_tests.log | head -40
