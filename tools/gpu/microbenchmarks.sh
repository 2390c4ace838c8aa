mkdir -p gpurun_out
for t in sttm_bench mma_issue_bench mma_loop_bench; do echo "=== tools/$t"; timeout 120 ./tools/$t; echo "rc=$?"; done > gpurun_out/r02_tcgen05_microbenchmarks.log 2>&1
tail -60 gpurun_out/r02_tcgen05_microbenchmarks.log
