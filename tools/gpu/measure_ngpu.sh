set -x; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${NG:-8}
timeout 300 $TR --nproc-per-node $N --master-port 29561 bench.py --gpus $N --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -2 gpurun_out/r02_bench_n$N.err
DP_CHECK_GRAPH=1 timeout 300 $TR --nproc-per-node $N --master-port 29562 tools/dp_check.py > gpurun_out/r02_dpcheck_n${N}_graph.log 2>&1; tail -3 gpurun_out/r02_dpcheck_n${N}_graph.log
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'))
PY
