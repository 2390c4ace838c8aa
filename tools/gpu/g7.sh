set -x; mkdir -p gpurun_out; rm -f gpurun_out/g7.log
for abl in ${ABLS:-0 1 2 4 8 16 31}; do
echo "== SDVAE_ABL=$abl" >> gpurun_out/g7.log
SDVAE_EXTRA_FLAGS="-DSDVAE_ABL=$abl" python craniofacialsd-vae_b200/build.py --force > /dev/null 2>&1
timeout 300 python tools/tile_check.py --levels 0 --B 1024 --skip-check --skip-old --only ${ONLY:-fwd} >> gpurun_out/g7.log 2>&1; echo "rc=$?" >> gpurun_out/g7.log
done
cat gpurun_out/g7.log
