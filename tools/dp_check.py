#!/usr/bin/env python
"""Data-parallel parity check (run under torchrun with N >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dp_check.py

Every rank runs one DP training step of the craniofacial model on its rows of the bs x bs swap
grid; rank 0 additionally runs the same step on ONE GPU with the whole grid and compares the
all-reduced gradient arena, the seven losses and the updated parameters.
DP_CHECK_GRAPH=1 runs the data-parallel step from a captured CUDA graph (as bench.py does)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from sdvae_b200 import fixtures as fx, losses
    from sdvae_b200.engine import StepConfig, TrainEngine

    tabs = fx.craniofacial_tables()
    bs = 2 * world
    cfg = StepConfig(batch_size=bs, lr=1e-3)
    lat = tabs.latent_regions(75)

    def make(pg):
        model = fx.build_model(tabs, 3, [32, 32, 32, 64], 75, False, True, 5, dev)     # same seed on every rank
        lt = losses.LaplacianTable.build(*tabs.lap, tabs.num_vertices[0], dev)
        return TrainEngine(model, lt, [r[1] for r in tabs.regions],
                           [lat[k] for k in tabs.region_keys()], cfg, process_group=pg,
                           use_graph=(pg is not None and os.environ.get('DP_CHECK_GRAPH') == '1'))

    rng = np.random.RandomState(1)
    x = torch.from_numpy(rng.randn(bs, tabs.num_vertices[0], 3).astype(np.float32)).to(dev)
    eps = torch.from_numpy(rng.randn(bs * bs, 75).astype(np.float32)).to(dev)
    region = 6

    eng = make(dist.group.WORLD)
    lo = eng.i0 * bs
    eng.set_fixed_eps(eps[lo:lo + eng.B])
    eng.load_batch(x)
    got = eng.step(region, sync_losses=True)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        ref = make(None)
        ref.set_fixed_eps(eps)
        ref.load_batch(x)
        want = ref.step(region, sync_losses=True)
        torch.cuda.synchronize()
        gerr = float((eng.flat_g - ref.flat_g).abs().max() / ref.flat_g.abs().max())
        perr = float((eng.flat_p - ref.flat_p).abs().max())
        lerr = max(abs(got[k] - want[k]) / max(abs(want[k]), 1e-12) for k in want if want[k] != 0)
        print("dp_check world=%d: grad normwise err %.2e, max |param diff| %.2e, loss rel err %.2e"
              % (world, gerr, perr, lerr))
        print("losses dp  ", got)
        print("losses 1gpu", want)
        ok = gerr < 5e-5 and lerr < 5e-5
    # ---- replicas stay bit-identical and draw DIFFERENT noise: three more steps with drawn eps ----
    eng.set_fixed_eps(None)
    for r in (2, 9, 2):
        eng.load_batch(x)
        eng.step(r, sync_losses=True)
    torch.cuda.synchronize()
    p0 = eng.flat_p.clone()
    dist.broadcast(p0, src=0)
    same_params = torch.equal(p0, eng.flat_p)
    e0 = eng.eps[0].clone()
    dist.broadcast(e0, src=0)
    noise_differs = rank == 0 or not torch.equal(e0, eng.eps[0])
    if rank == 0:
        print("replica parameters bit-identical after 4 steps: %s; eps differs across ranks: checked on ranks 1..%d"
              % (same_params, world - 1))
    ok = ok and same_params and noise_differs
    # release the captured step graphs (they hold NCCL work) before any further eager collective / teardown
    del eng
    import gc
    gc.collect()
    torch.cuda.synchronize()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("dp_check FAILED")
    if rank == 0:
        print("dp_check OK")


if __name__ == "__main__":
    main()
