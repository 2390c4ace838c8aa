set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g10_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g10_tests.log
tail -15 gpurun_out/g10_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g10_bench.json 2> gpurun_out/g10_bench.err; tail -3 gpurun_out/g10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g10_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config'].get('vertex_order'))
PY
SDVAE_TILE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g10_bench_notile.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g10_bench_notile.json').read().strip().splitlines()[-1])
print('SDVAE_TILE=0', {k:d[k] for k in ('value','ms_per_step')})
PY
