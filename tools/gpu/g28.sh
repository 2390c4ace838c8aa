set -x; mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
timeout 500 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -2 gpurun_out/r02_bench_reference_arm.err
timeout 300 python bench.py --bs 11 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_bs11.json 2>/dev/null
timeout 300 python tools/body_bench.py > gpurun_out/r02_body_n1.json 2> gpurun_out/r02_body_n1.err
timeout 900 python tools/microbench.py > gpurun_out/r02_microbench.md 2> gpurun_out/r02_microbench.err; tail -3 gpurun_out/r02_microbench.err
timeout 600 python tools/microbench.py --body > gpurun_out/r02_microbench_body.md 2> gpurun_out/r02_microbench_body.err; tail -3 gpurun_out/r02_microbench_body.err
timeout 600 python tools/infer_bench.py > gpurun_out/r02_infer_bench.md 2> gpurun_out/r02_infer_bench.err
CMD="python tools/tile_check.py --levels 0 --B 1024 --skip-check --skip-old --iters 1 --only fwd,dx,dw,out,bwo"
timeout 300 $CMD > gpurun_out/r02_tile_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gt_kernel|bt_kernel|pt_kernel|qt_kernel' -c 10 -o gpurun_out/r02_tile_full -f $CMD > gpurun_out/r02_tile_ncu.log 2>&1
tail -3 gpurun_out/r02_tile_ncu.log; cat gpurun_out/r02_tile_plain.log
BCMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $BCMD > gpurun_out/r02_bench_nograph.json 2> gpurun_out/r02_bench_nograph.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bs32.csv $BCMD > gpurun_out/r02_launches_ncu.log 2>&1
python - <<'PY'
import json
for f in ('r02_bench_n1','r02_bench_reference_arm','r02_bench_n1_bs11','r02_body_n1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','n_gpus','impl')}, (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('traffic'))
    except Exception as ex: print(f,'ERR',ex)
PY
