#!/usr/bin/env python
"""bench.py -- SD-VAE training-step throughput (meshes/s) on B200.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    torchrun ... bench.py --gpus N --steps K --warmup W           # data parallel, one rank per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own CPU path (baseline/_ref)

Workload (BASELINE.json configs[1]/[2]): craniofacial.yaml training step -- device-side feature
swap of ``bs`` synthetic N(0,1) template-shaped meshes into the ``bs x bs`` grid, forward,
MSE + 1e-4 KL + 0.5 latent-consistency + 0.1 Laplacian, backward, Adam -- at a GLOBAL batch of
bs^2 = 1024 meshes (bs = 32), sharded by grid rows over the ranks (strong scaling).

One JSON line on stdout (rank 0).  ``value`` = meshes/s with the un-swapped batch resident in
HBM; ``e2e`` = the same step through the public API with the batch in pinned HOST memory
(H2D copy + 28-byte loss read-back inside the timed region, one sync per step);
``roofline`` = the LONGEST kernel of the step timed alone with CUDA events (with the other two passes of
the same layer and the whole-step fraction beside it); ``cpu_baseline`` = the unmodified reference
(baseline/_ref, staged by tools/stage_reference.py; the oracle port only if it is absent) on this box's
host cores, bounded sample; ``reference_eager_cuda`` = the same reference modules run eagerly on the B200.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "sdvae_train_meshes_per_sec"
UNIT = "meshes/s"
CHANNELS = [32, 32, 32, 64]
LATENT = 75


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bs", type=int, default=32, help="swap-grid side; global batch = bs^2")
    ap.add_argument("--ref-bs", type=int, default=4, help="grid side of the CPU sample (reference yaml: 4)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="keep every contraction on the fp32-FMA kernels")
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(dev, bs, seed):
    """Tables (template vertex order; the engine renumbers its internal levels itself), the drop-in model on ``dev``
    (seeded xavier init), the synthetic un-swapped batch and the sequence of swapped regions.  No oracle on this
    path."""
    from sdvae_b200 import fixtures as fx
    tabs = fx.craniofacial_tables()
    model = fx.build_model(tabs, 3, CHANNELS, LATENT, False, True, seed, dev) if dev is not None else None
    rng = np.random.RandomState(seed)
    x = torch.from_numpy(rng.randn(bs, tabs.num_vertices[0], 3).astype(np.float32))
    regions = [int(r) for r in rng.randint(0, len(tabs.regions), 4096)]
    return tabs, model, x, regions


def cpu_step_rate(tabs, bs, steps, warmup, seed):
    """CPU-baseline leg: the reference's own training step (model.py + the loss methods of model_manager.py +
    SwapFeatures, baseline/refarm.py) on all host cores; only when the staged reference is absent, the oracle port
    (the one place bench.py executes oracle/).  Returns (meshes/s, s/step, kind)."""
    torch.set_num_threads(os.cpu_count() or 1)
    from baseline import refarm
    ref = refarm.find_ref()
    if ref is not None:
        rate, sec = refarm.time_reference_steps(ref, tabs, 'cpu', bs, steps, warmup, seed)
        return rate, sec, "reference"
    from oracle import sdvae_oracle as orc
    sp, dn, up = tabs.spiral_tensors(), tabs.down_tensors(), tabs.up_tensors()
    net = orc.Net(3, CHANNELS, LATENT, sp, dn, up, False, True)
    params = orc.xavier_params(net.param_shapes(), seed=seed)
    w = dict(kl=1e-4, lc=0.5, lap=0.1, eta1=0.5, eta2=0.5)
    trainer = orc.Trainer(net, params, tuple(torch.from_numpy(a) for a in tabs.lap), w, lr=1e-4)
    rng = np.random.RandomState(seed)
    x = torch.from_numpy(rng.randn(bs, tabs.num_vertices[0], 3).astype(np.float32))
    lat = tabs.latent_regions(LATENT)
    keys = tabs.region_keys()
    times = []
    for it in range(warmup + steps):
        r = int(rng.randint(0, len(keys)))
        t0 = time.perf_counter()
        xa = orc.swap_features(x, torch.from_numpy(tabs.regions[r][1]))
        trainer.step(xa, bs, lat[keys[r]])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    return bs * bs * len(times) / total, total / len(times), "port"


WORKLOAD = ("craniofacial.yaml train step (swap + fwd + MSE/KL/LC/Laplacian + bwd + Adam), "
            "V=17039, global batch %d")


def cpu_sample_text(kind, steps, bs):
    what = ("the UNMODIFIED reference (model.py, the loss methods of model_manager.py, SwapFeatures; baseline/_ref) "
            "on torch-CPU" if kind == "reference" else "oracle port of model.py/model_manager.py on torch-CPU "
            "(baseline/_ref absent)")
    return ("%d timed steps of the same training step on a BOUNDED SAMPLE of %d swapped meshes per step (bs=%d, the "
            "reference yaml's own batch size), fp32, all host threads, %s" % (steps, bs * bs, bs, what))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tabs, _, _, _ = build_problem(None, args.ref_bs, args.seed)
    rate, sec, kind = cpu_step_rate(tabs, args.ref_bs, args.steps, args.warmup, args.seed)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % (args.bs ** 2),
                   "sample": "%d swapped meshes per timed step (bs=%d): a bounded sample of the workload, not the "
                             "global batch" % (args.ref_bs ** 2, args.ref_bs)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, args.steps, args.ref_bs)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _time(run, reps):
    for _ in range(3):
        run()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / 1e3 / reps


def time_conv_kernels(eng, reps=10):
    """The three passes of the de4 SpiralConv (17039 x 288 x 32 per mesh: half of the model's FLOPs) -- forward+ELU,
    weight gradient, backward-to-input -- each timed alone with CUDA events on the launching stream, on the engine's
    own buffers (1024 x 17039 x 32 floats each: far beyond L2).  Returns {pass: (seconds, kernel name)} and the
    algorithmic bytes / flops of one pass."""
    from sdvae_b200 import cabi
    L, V, C, B, S = eng.L, eng.V, eng.C, eng.B, eng.S
    layer = eng.model.de_layers[L].conv.layer
    cin, cout = eng.cin_de[0], C[1]
    eng._pack_tc()
    out = {}
    ef, eb = eng.tc.get(('f', 'de0')), eng.tc.get(('b', 'de0'))
    tiled = ef is not None and ef['tile'] is not None
    out['fwd'] = (_time(lambda: eng._conv(eng.u[0], eng.full[0], layer, eng.d[0], cabi.ACT_ELU, B, V[0], cin, cout, 'de0'), reps),
                  "gt_kernel<uniform> (tcgen05 3xTF32, tile-staged)" if tiled else
                  ("gc_umma_kernel<32,32,uniform> (tcgen05 3xTF32)" if ef is not None else "gc_tile_kernel (fp32 FMA)"))
    out['dW'] = (_time(lambda: eng._bwd_w(eng.u[0], eng.full[0], eng.dd[0], layer, B, V[0], cin, cout), reps),
                 "bt_kernel (tcgen05 3xTF32, tile-staged)" if tiled and eng.full[0].tile_fwd() is not None else
                 ("bw_umma_kernel (tcgen05 3xTF32)" if eng.use_tc else "bw_outer_kernel (fp32 FMA)"))
    if eb is not None and eb['tile'] is not None:
        run = lambda: cabi.spiralconv_bwd_x_tile(eng.dd[0], eb['tile'], eb['parts'][0][2], None, eng.du[0], B, V[0], V[0], S[0], cout, cin)
        name = "gt_kernel<ragged> (tcgen05 3xTF32, tile-staged)"
    elif eb is not None:
        run = lambda: cabi.spiralconv_bwd_x_tc(eng.dd[0], eb['plan'], eb['parts'][0][2], None, eng.du[0], B, V[0], V[0], S[0], cout, cin, 0)
        name = "gc_umma_kernel<32,32,ragged> (tcgen05 3xTF32)"
    else:
        run, name = None, None
    if run is not None:
        out['dx'] = (_time(run, reps), name)
    alg_bytes = 4.0 * B * V[0] * (cin + cout) + 4.0 * V[0] * S[0] + 4.0 * (S[0] * cin * cout + cout)
    flops = 2.0 * B * V[0] * S[0] * cin * cout
    return out, alg_bytes, flops


def time_out_layer(eng, reps=10):
    """The two passes of the 32 -> 3 output layer (forward; fused backward = dx with the previous ELU', dW, db), each
    timed alone with CUDA events.  Returns {pass: (seconds, kernel name, algorithmic bytes)} (the forward
    moves 4*(32 + 3) floats per vertex, the backward reads dy and x and writes dx: 4*(3 + 32 + 32))."""
    from sdvae_b200 import cabi
    L, V, C, B, S = eng.L, eng.V, eng.C, eng.B, eng.S
    lay = eng.model.de_layers[L + 1].layer
    out = {}
    if eng.narrow_out_tile is not None:
        out['fwd'] = (_time(lambda: cabi.narrow_out_fwd_tc(eng.d[0], eng.narrow_out_tile, lay.weight.data, lay.bias.data,
                                                           eng.recon, B, V[0], V[0], S[0], C[1], C[0]), reps),
                      "pt_kernel (tcgen05 3xTF32, project-then-gather)", 4.0 * B * V[0] * (C[1] + C[0]))
    if eng.narrow_out_tile_bwd is not None:
        out['bwd'] = (_time(lambda: cabi.narrow_out_bwd_tc(eng.drecon, eng.d[0], eng.narrow_out_tile_bwd, lay.weight.data,
                                                           eng.dd[0], eng.g(lay.weight), eng.g(lay.bias), eng.narrow_out_tc_ws,
                                                           B, V[0], V[0], S[0], C[1], C[0], True), reps),
                      "qt_kernel (tcgen05 3xTF32, fused gather-then-project: dx, dW, db)", 4.0 * B * V[0] * (C[0] + 2 * C[1]))
    return out


def time_pool_kernel(eng, reps=10):
    """The finest up-sampling Pool forward ([B, 4260, 32] -> [B, 17039, 32], the largest memory-bound Pool of
    the step), timed alone with CUDA events on the launching stream; inputs / outputs exceed L2 at the bench batch."""
    from sdvae_b200 import cabi
    V, B, C = eng.V, eng.B, eng.cin_de[0]
    up = eng.up[0]
    src = eng.d[1] if eng.L > 1 else eng.h
    run = lambda: cabi.pool_fwd(src, up, eng.u[0], B, V[1], C)
    for _ in range(3):
        run()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    sec = s.elapsed_time(e) / 1e3 / reps
    alg_bytes = 4.0 * B * C * (V[1] + V[0]) + 8.0 * up.width * V[0]
    return sec, alg_bytes


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sdvae_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    from sdvae_b200 import cabi, losses
    from sdvae_b200.engine import StepConfig, TrainEngine

    bs = args.bs
    tabs, model, x_host, regions = build_problem(dev, bs, args.seed)
    cfg = StepConfig(batch_size=bs)
    lt = losses.LaplacianTable.build(*tabs.lap, tabs.num_vertices[0], dev)
    lat = tabs.latent_regions(LATENT)
    eng = TrainEngine(model, lt, [r[1] for r in tabs.regions], [lat[k] for k in tabs.region_keys()],
                      cfg, process_group=pg, use_graph=not args.no_graph, use_tc=not args.no_tc)
    x_pin = x_host.pin_memory()
    eng.load_batch(x_pin)
    torch.cuda.synchronize()
    K, W = args.steps, args.warmup
    seq = regions[:W + K]
    eng.prepare(sorted(set(regions[:2 * (W + K)])))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------
    for r in seq[:W]:
        eng.step(r)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = cabi.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for r in seq[W:W + K]:
        eng.step(r)
    e.record()
    barrier()
    launches = cabi.launch_count() - n0
    ms = max_over_ranks(s.elapsed_time(e))
    clocks = sampler.stop() if sampler else None
    meshes = bs * bs * K
    value = meshes / (ms / 1e3)

    # ---- end to end: pinned host batch in, 7 loss scalars out, every step --------------
    # Every step's un-swapped batch is copied from pinned host memory inside the timed region and every step's losses
    # are read back to the host and waited for.  As in any training loop with an input pipeline, the copy of batch k+1
    # is issued while step k runs (the engine lands it in a staging buffer on its copy stream, TrainEngine.load_batch)
    # and the losses of step k are consumed after step k+1 has been launched (logging one step late).
    seq2 = regions[W + K:2 * (W + K)]

    def e2e_steps(rs):
        eng.load_batch(x_pin)
        for i, r in enumerate(rs):
            eng.step(r)                                   # consumes the staged batch, launches step i + its loss read-back
            if i + 1 < len(rs):
                eng.load_batch(x_pin)                     # H2D of the next batch overlaps this step
            if i > 0:
                eng.wait_losses(lag=1)                    # losses of step i-1 (D2H done long ago)
        eng.wait_losses()                                 # ... and of the last step

    e2e_steps(seq2[:W])
    barrier()
    t0 = time.perf_counter()
    s.record()
    e2e_steps(seq2[W:W + K])
    e.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(s.elapsed_time(e), wall_ms))
    e2e_value = meshes / (ms_e2e / 1e3)
    last = eng.loss_dict()

    if rank != 0:
        return
    # ---- roofline: the LONGEST kernel of the step (rank 0, timed alone), the layer's other passes, the whole step ----
    hbm_peak, peak_src = peaks()
    passes, alg_bytes, flops = time_conv_kernels(eng)
    worst = max(passes, key=lambda k: passes[k][0])
    ksec, kname = passes[worst]
    achieved = alg_bytes / ksec / 1e9
    # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of an `ncu --set full` capture of the same
    # kernel at 1024 meshes (profiles/r02_traffic.json names the capture files and the commit), scaled to this run's
    # mesh count; null when no capture of this kernel is on file.
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(kname.split(" ")[0] + ":" + worst)
        if t:
            traffic = float(t["dram_bytes_at_1024_meshes"]) * eng.B / 1024.0
    # whole step: forward + backward ~ 3 x the forward's fused-op bytes (19.79 MB per mesh forward)
    step_bytes = 3.0 * 19.79e6 * eng.B
    roofline = {"bound": "hbm",
                "kernel": "%s de4 SpiralConv %s [%d x 17039 x 288 x 32]" % (kname, worst, eng.B),
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "kernel_ms": ksec * 1e3, "effective_tflops": flops / ksec / 1e12,
                "selection": "longest launch of the step (the three passes of the de4 layer are its three longest kernels)",
                "de4_passes": {k: {"kernel": v[1], "kernel_ms": v[0] * 1e3, "achieved": alg_bytes / v[0] / 1e9,
                                   "frac": alg_bytes / v[0] / 1e9 / hbm_peak} for k, v in passes.items()},
                "whole_step": {"algorithmic_bytes": step_bytes, "ms_per_step": ms / K,
                               "achieved": step_bytes / (ms / K / 1e3) / 1e9,
                               "frac": step_bytes / (ms / K / 1e3) / 1e9 / hbm_peak,
                               "note": "3 x 19.79 MB per mesh (forward + two backward passes of every fused op) over this "
                                       "rank's step time"},
                "note": ("error-compensated 3xTF32 on tcgen05 (fp32-level parity).  Not HBM-bound: the tile's distinct "
                         "source rows are staged once in shared memory (DRAM traffic = algorithmic bytes, `traffic`); the "
                         "weight gradient is bound by instruction issue (64-72 % of the issue slots: one LDS.32 per element "
                         "in the transposing splitter) and, like the forward / input-gradient kernels, by the TMEM stores "
                         "of the gathered, hi/lo-split A operand (tcgen05.st ~155 clk per 4 KB and warp + ~300 clk "
                         "tcgen05.wait::st, tools/sttm_bench.cu; profiles/r02_tile_kernel.md)")}
    out_passes = time_out_layer(eng)
    roofline["out_layer_passes"] = {k: {"kernel": v[1], "kernel_ms": v[0] * 1e3, "algorithmic_bytes": v[2],
                                        "achieved": v[2] / v[0] / 1e9, "frac": v[2] / v[0] / 1e9 / hbm_peak}
                                    for k, v in out_passes.items()}
    psec, palg = time_pool_kernel(eng)
    use_graph, renumbered, has_tc, meshes_per_gpu, n_params = bool(eng.use_graph), bool(eng.renumber), bool(eng.tc), eng.B, eng.n_params
    pool_B, pool_V = eng.B, (eng.V[1], eng.V[0])
    pool_roofline = {"bound": "hbm", "kernel": "pool_ell_fwd_staged_kernel: Pool up-sampling fwd [%d x %d -> %d x 32]"
                                               % (pool_B, pool_V[0], pool_V[1]),
                     "achieved": palg / psec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": palg / psec / 1e9 / hbm_peak, "algorithmic_bytes": palg, "kernel_ms": psec * 1e3,
                     # ncu --set full of this launch at 1024 meshes (profiles/r01_pool_staged_full_metrics.txt)
                     "traffic": (0.583874e9 + 2.175260e9) * pool_B / 1024.0, "peak_source": peak_src,
                     "note": "distinct source rows of each 128-row tile staged in shared memory (cp.async ring); "
                             "bit-identical to the reference's storage-order arithmetic"}
    cpu = None
    ref_cuda = None
    if not args.no_cpu_baseline and world == 1:      # reported baselines: rank 0 at N = 1 only
        rate, sec, kind = cpu_step_rate(tabs, args.ref_bs, 5, 2, args.seed)
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
               "sample": cpu_sample_text(kind, 5, args.ref_bs), "ms_per_step": sec * 1e3}
        # the kernel-for-kernel comparator: the same unmodified reference modules, eager ATen / cuBLAS
        # kernels on this B200, fp32 (TF32 off, torch default), 256 swapped meshes per step (the reference keeps every
        # materialised gather for backward: ~75 MB per mesh)
        from baseline import refarm
        ref = refarm.find_ref()
        if ref is not None:
            del eng
            torch.cuda.empty_cache()
            try:
                r2, s2 = refarm.time_reference_steps(ref, tabs, dev, 16, 3, 2, args.seed)
                ref_cuda = {"value": r2, "unit": UNIT, "ms_per_step": s2 * 1e3, "meshes_per_step": 256,
                            "what": "reference model.py + lifted loss methods + SwapFeatures (CPU collate, as in the "
                                    "reference's DataLoader), torch eager on cuda, 3 timed steps"}
            except Exception as ex:      # never lose the bench line to the comparator
                ref_cuda = {"unavailable": repr(ex)[:200]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (bs * bs),
                   "global_batch": bs * bs, "grid": "%dx%d" % (bs, bs), "meshes_per_gpu": meshes_per_gpu,
                   "parallelism": "dp%d (swap-grid rows)" % world,
                   "l2": "per-step working set (%.1f GB/GPU) exceeds the 126 MB L2" % (meshes_per_gpu * 12.0e6 / 1e9),
                   "cuda_graph": use_graph, "vertex_order": "template outside the engine; internal levels patch-wise" if renumbered else "template",
                   "contractions": ("tcgen05 3xTF32 for every 32/64-channel SpiralConv pass and for the 32->3 output layer "
                                    "(project-then-gather forward, fused gather-then-project backward); the 3->32 first "
                                    "block on the fp32 FMA units from shared-memory-resident meshes")
                                   if has_tc else "fp32 FMA"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": int(x_pin.numel() * 4), "d2h_bytes_per_step": 32},
        "roofline": roofline, "pool_roofline": pool_roofline, "cpu_baseline": cpu, "reference_eager_cuda": ref_cuda,
        "losses_last_step": last, "params": n_params,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
