set -x; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for i in 1 2; do
timeout 300 $TR --nproc-per-node 2 --master-port 2957$i bench.py --gpus 2 --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/r02_bench_n2_run$i.json 2> gpurun_out/r02_bench_n2_run$i.err; echo "rc=$?"; tail -1 gpurun_out/r02_bench_n2_run$i.err | cut -c1-200
done
python - <<'PY'
import json
for i in (1,2):
    try:
        d=json.loads(open('gpurun_out/r02_bench_n2_run%d.json'%i).read().strip().splitlines()[-1])
        print(i,{k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'))
    except Exception as ex: print(i,'ERR',ex)
PY
