set -x; mkdir -p gpurun_out
timeout 600 python tools/tile_check.py --levels 0 --B 1024 > gpurun_out/g2_tile_check.log 2>&1; echo "rc=$?" >> gpurun_out/g2_tile_check.log
tail -40 gpurun_out/g2_tile_check.log
