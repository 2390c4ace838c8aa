set -x; mkdir -p gpurun_out
timeout 200 python tools/tile_check.py --levels 0 --B 2 --only out,bwo --iters 1 > gpurun_out/g30_plain.log 2>&1; echo "rc=$?" >> gpurun_out/g30_plain.log; cat gpurun_out/g30_plain.log
timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python tools/tile_check.py --levels 0 --B 2 --only out,bwo --iters 1 > gpurun_out/g30_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/g30_memcheck.log
grep -v "^  out\|^level" gpurun_out/g30_memcheck.log | head -60
