set -x; mkdir -p gpurun_out
timeout 900 python tools/microbench.py > gpurun_out/r02_microbench.md 2> gpurun_out/r02_microbench.err; echo "rc=$?"; tail -2 gpurun_out/r02_microbench.err
timeout 600 python tools/microbench.py --body > gpurun_out/r02_microbench_body.md 2> gpurun_out/r02_microbench_body.err; echo "rc=$?"; tail -2 gpurun_out/r02_microbench_body.err
CMD="python tools/tile_check.py --levels 0 --B 1024 --skip-check --skip-old --iters 1 --only fwd,dx,dw,out,bwo"
timeout 300 $CMD > gpurun_out/r02_tile_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gt_kernel|bt_kernel|pt_kernel|qt_kernel' -c 10 -o gpurun_out/r02_tile_full -f $CMD > gpurun_out/r02_tile_ncu.log 2>&1
tail -4 gpurun_out/r02_tile_ncu.log
