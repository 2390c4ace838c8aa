set -x; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g24_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g24_tests.log; tail -3 gpurun_out/g24_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g24_bench_n1.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g24_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, (d.get('e2e') or {}).get('value'))
PY
