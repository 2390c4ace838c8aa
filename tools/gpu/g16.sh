set -x; mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
timeout 200 python tools/microbench.py --batches 1,16 --iters 3 > gpurun_out/r02_microbench_smoke.md 2> gpurun_out/r02_microbench_smoke.err; tail -3 gpurun_out/r02_microbench_smoke.err
timeout 900 python tools/microbench.py > gpurun_out/r02_microbench.md 2> gpurun_out/r02_microbench.err; tail -3 gpurun_out/r02_microbench.err
timeout 600 python tools/microbench.py --body > gpurun_out/r02_microbench_body.md 2> gpurun_out/r02_microbench_body.err; tail -3 gpurun_out/r02_microbench_body.err
timeout 600 python tools/infer_bench.py > gpurun_out/r02_infer_bench.md 2> gpurun_out/r02_infer_bench.err; tail -3 gpurun_out/r02_infer_bench.err; cat gpurun_out/r02_infer_bench.md
CMD="python tools/tile_check.py --levels 0 --B 1024 --skip-check --skip-old --iters 1"
timeout 300 $CMD > gpurun_out/r02_tile_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gt_kernel|bt_kernel' -c 6 -o gpurun_out/r02_tile_full -f $CMD > gpurun_out/r02_tile_ncu.log 2>&1
tail -3 gpurun_out/r02_tile_ncu.log
BCMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $BCMD > gpurun_out/r02_bench_nograph.json 2> gpurun_out/r02_bench_nograph.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bs32.csv $BCMD > gpurun_out/r02_launches_ncu.log 2>&1
tail -2 gpurun_out/r02_launches_ncu.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d.get('e2e'), d.get('roofline'))
PY
