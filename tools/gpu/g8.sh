set -x; mkdir -p gpurun_out
timeout 600 python tools/tile_check.py --levels ${LEVELS:-0} --B 1024 ${EXTRA:-} > gpurun_out/g8_tile_check.log 2>&1; echo "rc=$?" >> gpurun_out/g8_tile_check.log
cat gpurun_out/g8_tile_check.log
