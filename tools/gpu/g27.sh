set -x; mkdir -p gpurun_out
CMD="python tools/tile_check.py --levels 0 --B 512 --only bwo --skip-check --iters 1"
timeout 200 $CMD > gpurun_out/g27_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qt_kernel -s 1 -c 1 -o gpurun_out/g27_qt -f $CMD > gpurun_out/g27_ncu.log 2>&1
tail -3 gpurun_out/g27_ncu.log
