mkdir -p gpurun_out
timeout 300 python tools/infer_bench.py > gpurun_out/r02_infer_bench.md 2> gpurun_out/r02_infer_bench.err; echo "rc=$?"; cat gpurun_out/r02_infer_bench.md; tail -3 gpurun_out/r02_infer_bench.err
