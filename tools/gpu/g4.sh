set -x; mkdir -p gpurun_out
CMD="python tools/tile_check.py --levels 0 --B 256 --skip-check --skip-old --only ${ONLY:-fwd} --iters 2"
$CMD > gpurun_out/g4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gt_kernel -s 1 -c 1 -o gpurun_out/g4_gt_${ONLY:-fwd} -f $CMD > gpurun_out/g4_ncu.log 2>&1
tail -5 gpurun_out/g4_ncu.log
