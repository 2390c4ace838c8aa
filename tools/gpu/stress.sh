set -x; mkdir -p gpurun_out
for bs in 23 16 11 32; do
timeout 300 python bench.py --bs $bs --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/stress_bs$bs.json 2> gpurun_out/stress_bs$bs.err; echo "bs=$bs rc=$?"
done
python - <<'PY'
import json
for bs in (23,16,11,32):
    try:
        d=json.loads(open('gpurun_out/stress_bs%d.json'%bs).read().strip().splitlines()[-1])
        print(bs,{k:d.get(k) for k in ('value','ms_per_step')}, d['losses_last_step']['tot'], d['clocks'])
    except Exception as ex: print(bs,'ERR',ex)
PY
