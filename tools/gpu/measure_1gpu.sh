set -x; mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --bs 11 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_bs11.json 2>/dev/null
timeout 300 python tools/body_bench.py > gpurun_out/r02_body_n1.json 2> gpurun_out/r02_body_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
BCMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $BCMD > gpurun_out/r02_bench_nograph.json 2> gpurun_out/r02_bench_nograph.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bs32.csv $BCMD > gpurun_out/r02_launches_ncu.log 2>&1
python - <<'PY'
import json
for f in ('r02_bench_n1','r02_bench_n1_bs11','r02_body_n1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f,{k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, (d.get('e2e') or {}).get('value'), d.get('cpu_baseline',{}) and d['cpu_baseline'].get('value'), (d.get('reference_eager_cuda') or {}).get('value'))
        r=d.get('roofline') or {}
        print('   ', r.get('frac'), r.get('traffic'), {k:round(v['frac'],3) for k,v in (r.get('de4_passes') or {}).items()}, {k:round(v['frac'],3) for k,v in (r.get('out_layer_passes') or {}).items()}, (r.get('whole_step') or {}).get('frac'))
    except Exception as ex: print(f,'ERR',ex)
PY
