set -x; mkdir -p gpurun_out
export SDVAE_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -m gpu -k experimental -x -q > gpurun_out/g1_exp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g1_exp_tests.log
tail -5 gpurun_out/g1_exp_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g1_bench_base.json 2> gpurun_out/g1_bench_base.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --renumber > gpurun_out/g1_bench_renum.json 2> gpurun_out/g1_bench_renum.err
SDVAE_STAGED_FWD=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --renumber > gpurun_out/g1_bench_renum_staged.json 2> gpurun_out/g1_bench_renum_staged.err
SDVAE_STAGED_FWD=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g1_bench_staged.json 2> gpurun_out/g1_bench_staged.err
cat gpurun_out/g1_bench_*.json
