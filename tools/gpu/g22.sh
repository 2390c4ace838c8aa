set -x; mkdir -p gpurun_out
timeout 300 python tools/tile_check.py --levels 0 --B 1024 --only out --iters 5 > gpurun_out/g22_out.log 2>&1; echo "rc=$?" >> gpurun_out/g22_out.log; cat gpurun_out/g22_out.log
