set -x; mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/microbench.py --batches 2 --iters 3 > gpurun_out/g31_mb2.md 2> gpurun_out/g31_mb2.err; echo "rc=$?"; tail -3 gpurun_out/g31_mb2.md; grep -n "File\|Error" gpurun_out/g31_mb2.err | tail -8
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/microbench.py --body --batches 8 --iters 3 > gpurun_out/g31_mbb8.md 2> gpurun_out/g31_mbb8.err; echo "rc=$?"; tail -3 gpurun_out/g31_mbb8.md; grep -n "File\|Error" gpurun_out/g31_mbb8.err | tail -8
