set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/g14_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g14_tests.log; tail -3 gpurun_out/g14_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/g14_dpcheck_n2.log 2>&1; tail -5 gpurun_out/g14_dpcheck_n2.log
DP_CHECK_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/g14_dpcheck_n2_graph.log 2>&1; tail -5 gpurun_out/g14_dpcheck_n2_graph.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g14_bench_n2.json 2> gpurun_out/g14_bench_n2.err; tail -2 gpurun_out/g14_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/g14_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g14_bench_n1.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/g14_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
"
