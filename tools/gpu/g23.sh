set -x; mkdir -p gpurun_out
for m in 1 2 4 7; do echo "== ABL_OUT=$m"; timeout 200 python tools/tile_check.py --levels 0 --B 1024 --only out --iters 5 --skip-check --lib variants/lib_out_abl$m.so 2>&1 | grep "project-then"; done > gpurun_out/g23_abl.log 2>&1
cat gpurun_out/g23_abl.log
