set -x; mkdir -p gpurun_out
BCMD="python bench.py --bs 11 --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
timeout 300 $BCMD > gpurun_out/g18_bench_nograph.json 2> gpurun_out/g18_bench_nograph.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bs11.csv $BCMD > gpurun_out/g18_ncu.log 2>&1
tail -2 gpurun_out/g18_ncu.log
timeout 300 python bench.py --bs 11 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/g18_bench_bs11.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g18_bench_bs11.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d.get('e2e'))
PY
