set -x; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --bs 16 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/g11_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bs16.csv $CMD > gpurun_out/g11_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_bs16.csv > gpurun_out/r02_launches_bs16_summary.md; cat gpurun_out/r02_launches_bs16_summary.md
