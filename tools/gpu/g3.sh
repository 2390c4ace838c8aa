set -x; mkdir -p gpurun_out
SDVAE_TUNING=1 python craniofacialsd-vae_b200/build.py --force > /dev/null 2>&1
for dbg in ${DBGS:-32}; do
echo "== SDVAE_DBG=$dbg" >> gpurun_out/g3_prof.log
SDVAE_DBG=$dbg timeout 300 python tools/tile_check.py --levels 0 --B ${B:-1024} --skip-check --skip-old ${EXTRA:-} >> gpurun_out/g3_prof.log 2>&1; echo "rc=$?" >> gpurun_out/g3_prof.log
done
cat gpurun_out/g3_prof.log
