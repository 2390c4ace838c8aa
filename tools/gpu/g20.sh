set -x; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
SDVAE_DP_OVERLAP=0 DP_CHECK_GRAPH=1 timeout 300 $TR --master-port 29551 tools/dp_check.py > gpurun_out/g20_dpcheck_single.log 2>&1; tail -4 gpurun_out/g20_dpcheck_single.log
SDVAE_DP_OVERLAP=0 timeout 300 $TR --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g20_single_n2.json 2> gpurun_out/g20_single_n2.err; tail -2 gpurun_out/g20_single_n2.err
SDVAE_DP_OVERLAP=1 timeout 300 $TR --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g20_overlap_n2.json 2> gpurun_out/g20_overlap_n2.err
python - <<'PY'
import json
for f in ('single_n2','overlap_n2'):
    try:
        d=json.loads(open('gpurun_out/g20_%s.json'%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'))
    except Exception as ex: print(f,'ERR',ex)
PY
