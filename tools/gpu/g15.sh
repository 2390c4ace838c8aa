set -x; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
DP_CHECK_GRAPH=1 timeout 300 $TR --nproc-per-node 8 --master-port 29521 tools/dp_check.py > gpurun_out/r02_dpcheck_n8_graph.log 2>&1; tail -4 gpurun_out/r02_dpcheck_n8_graph.log
timeout 300 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -2 gpurun_out/r02_bench_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
timeout 300 $TR --nproc-per-node 8 --master-port 29524 tools/body_bench.py --steps 30 --warmup 5 > gpurun_out/r02_body_n8.json 2> gpurun_out/r02_body_n8.err; tail -2 gpurun_out/r02_body_n8.err
timeout 300 python tools/body_bench.py > gpurun_out/r02_body_n1.json 2> gpurun_out/r02_body_n1.err
timeout 300 python bench.py --bs 11 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_bs11.json 2>/dev/null
python - <<'PY'
import json
for f in ('r02_bench_n8','r02_bench_n4','r02_body_n8','r02_body_n1','r02_bench_n1_bs11'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'))
    except Exception as ex: print(f,'ERR',ex)
PY
