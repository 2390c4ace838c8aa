set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g13_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g13_tests.log
tail -4 gpurun_out/g13_tests.log
timeout 600 python bench.py > gpurun_out/g13_bench.json 2> gpurun_out/g13_bench.err; tail -3 gpurun_out/g13_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g13_bench_ref.json 2> gpurun_out/g13_bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/g13_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','e2e','clocks','cpu_baseline','reference_eager_cuda')}, indent=1))
print(json.dumps(d['roofline'], indent=1)[:2500])
"
cat gpurun_out/g13_bench_ref.json | cut -c1-400
python __graft_entry__.py smoke 2>&1 | tail -2
