mkdir -p gpurun_out
run() { timeout 200 python tools/microbench.py --iters 3 "$@" > gpurun_out/g33_tmp.md 2> gpurun_out/g33_tmp.err; echo "rc=$? :: $*  :: last row: $(tail -1 gpurun_out/g33_tmp.md | cut -c1-90)"; }
run --batches 1,2 --only pool
run --batches 1,2 --layer de5
run --batches 1,2 --layer de4
run --batches 1,2 --layer en
run --batches 1,2 --layer de3
run --batches 1,2 --layer de2
run --batches 1,2 --layer de1
run --batches 2
run --batches 1,2
