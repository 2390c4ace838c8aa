set -x; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g21_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g21_tests.log; tail -3 gpurun_out/g21_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g21_bench_n1.json 2>/dev/null
timeout 300 python bench.py --bs 11 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/g21_bench_bs11.json 2>/dev/null
python - <<'PY'
import json
for f in ('g21_bench_n1','g21_bench_bs11'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f,{k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'), d['roofline']['de4_passes'])
PY
