set -x; mkdir -p gpurun_out; rm -f gpurun_out/g9.log
for k in ${WARPS:-1 6 11 16}; do
echo "== SDVAE_PROF=$k" >> gpurun_out/g9.log
SDVAE_EXTRA_FLAGS="-DSDVAE_PROF=$k" python craniofacialsd-vae_b200/build.py --force > /dev/null 2>&1
SDVAE_PROF=$k timeout 300 python tools/tile_check.py --levels 0 --B 1024 --skip-check --skip-old --only ${ONLY:-fwd} >> gpurun_out/g9.log 2>&1; echo "rc=$?" >> gpurun_out/g9.log
done
cat gpurun_out/g9.log
