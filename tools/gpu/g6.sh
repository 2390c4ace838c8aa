set -x; mkdir -p gpurun_out; rm -f gpurun_out/g6.log
for ns in ${SETS:-4 5 6}; do
echo "== SDVAE_TILE_SETS=$ns" >> gpurun_out/g6.log
SDVAE_TILE_SETS=$ns timeout 600 python tools/tile_check.py --levels ${LEVELS:-0} --B 1024 --skip-old ${EXTRA:-} >> gpurun_out/g6.log 2>&1; echo "rc=$?" >> gpurun_out/g6.log
done
cat gpurun_out/g6.log
