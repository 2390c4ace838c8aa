set -x; mkdir -p gpurun_out
timeout 300 python tools/tile_check.py --levels 0 --B 1024 --skip-check --skip-old --iters 5 --only out,bwo > gpurun_out/g36_tile.log 2>&1; cat gpurun_out/g36_tile.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g36_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g36_tests.log; tail -3 gpurun_out/g36_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g36_bench_n1.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g36_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, (d.get('e2e') or {}).get('value'), d['roofline'].get('out_layer_passes'))
PY
