set -x; mkdir -p gpurun_out
timeout 120 python tools/tile_check.py --levels 0 --B 1024 --only bwo --iters 5 > gpurun_out/g25_bwo.log 2>&1; echo "rc=$?" >> gpurun_out/g25_bwo.log; cat gpurun_out/g25_bwo.log
