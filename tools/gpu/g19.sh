set -x; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $TR --master-port $PORT bench.py --gpus 8 --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/g19_$name.json 2> gpurun_out/g19_$name.err
  PORT=$((PORT+1))
}
PORT=29541
run overlap SDVAE_DP_OVERLAP=1
run single SDVAE_DP_OVERLAP=0
run overlap_cta4 SDVAE_DP_OVERLAP=1 NCCL_MAX_CTAS=4
run single_cta4 SDVAE_DP_OVERLAP=0 NCCL_MAX_CTAS=4
python - <<'PY'
import json
for f in ('overlap','single','overlap_cta4','single_cta4'):
    try:
        d=json.loads(open('gpurun_out/g19_%s.json'%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'))
    except Exception as ex: print(f,'ERR',ex)
PY
