set -x; mkdir -p gpurun_out
timeout 900 python tools/microbench.py > gpurun_out/r02_microbench.md 2> gpurun_out/r02_microbench.err; echo "rc=$?"; tail -2 gpurun_out/r02_microbench.err
timeout 600 python tools/microbench.py --body > gpurun_out/r02_microbench_body.md 2> gpurun_out/r02_microbench_body.err; echo "rc=$?"; tail -2 gpurun_out/r02_microbench_body.err
timeout 200 python tools/tile_check.py --levels 0 --B 1024 --only bwo --iters 5 > gpurun_out/g29_bwo.log 2>&1; echo "rc=$?" >> gpurun_out/g29_bwo.log; cat gpurun_out/g29_bwo.log
