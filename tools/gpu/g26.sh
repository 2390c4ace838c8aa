set -x; mkdir -p gpurun_out
for m in 1 2 3 4 8 16 32 63; do echo "== ABL_QT=$m"; timeout 200 python tools/tile_check.py --levels 0 --B 1024 --only bwo --iters 5 --skip-check --lib variants/lib_qt_abl$m.so 2>&1 | grep "gather-then"; done > gpurun_out/g26_abl.log 2>&1
cat gpurun_out/g26_abl.log
