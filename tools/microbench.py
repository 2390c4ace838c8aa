#!/usr/bin/env python
"""SpiralConv + Pool micro-benchmark sweep (BASELINE.json configs[4]).

Every encoder/decoder layer of craniofacial.yaml x {forward, backward-to-input, weight gradient} and every
Pool level x {forward, backward}, at several batch sizes, through the C ABI, timed with CUDA events on the
launching stream.  Each timed iteration works on a different one of several input/output buffer sets
whose total size exceeds the 126 MB L2 (so no iteration finds its inputs cached).  Reports milliseconds,
the ALGORITHMIC bytes per launch (4*(Vin*Cin + Vout*Cout) per mesh for a conv,
4*C*(rows_read + Vout) for a pool), the achieved GB/s against MEASURED_PEAKS.json's HBM figure, and the
effective TFLOP/s.  Output: a Markdown table on stdout (committed under profiles/).

usage: python tools/microbench.py [--batches 1,16,256,1024] [--iters 10]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = 'cuda:0'


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


def timeit(fn, sets, iters):
    """fn(k) runs the launch on buffer set k.  Returns average ms per launch."""
    for k in range(min(3, sets)):
        fn(k % sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % sets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def nsets(bytes_per_set):
    return int(max(2, min(8, (300 << 20) // max(1, bytes_per_set) + 1)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', default='1,2,4,8,16,32,64,128,256,512,1024,2048,4096',
                    help='BASELINE.json configs[4]: batch 1-4096')
    ap.add_argument('--body', action='store_true', help='body.yaml layer set on the synthetic 6890-vertex template')
    ap.add_argument('--template-order', action='store_true',
                    help="keep the template's own vertex order (the engine renumbers its internal levels patch-wise)")
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--layer', default='', help='substring of the conv layer names to run')
    ap.add_argument('--only', default='', help="'pool' or 'conv': run only that half of the sweep")
    args = ap.parse_args()
    from sdvae_b200 import cabi, fixtures as fx
    from sdvae_b200.tables import identity_plan, pool_table, restricted_spiral_table, spiral_table
    cabi.load()
    if args.body:
        tabs = fx.synthetic_tables(6890, 3, seq_length=9, n_regions=11, seed=0, name='body-synthetic')
    else:
        tabs = fx.craniofacial_tables()
    if not args.template_order:
        tabs = tabs.renumbered(128)[0]
    sp = [s.to(DEV) for s in tabs.spiral_tensors()]
    dn = [d.to(DEV) for d in tabs.down_tensors()]
    up = [u.to(DEV) for u in tabs.up_tensors()]
    V = tabs.num_vertices
    peak, src = hbm_peak()
    f = lambda *shape: torch.randn(shape, device=DEV, dtype=torch.float32)
    print('# SpiralConv / Pool micro-benchmark (%s layers, %s vertex order, 1 x B200)\n'
          % ('body.yaml, synthetic 6890-vertex template' if args.body else 'craniofacial.yaml',
             'template' if args.template_order else 'patch-wise (as inside TrainEngine)'))
    print('HBM peak %.1f GB/s (%s); algorithmic bytes per SURVEY.md 8(d); inputs rotate through buffer sets > L2.\n' % (peak, src))
    print('| op | layer | B | path | ms | alg MB | GB/s | frac of HBM peak | TFLOP/s |')
    print('|---|---|---|---|---|---|---|---|---|')

    def row(op, layer, B, path, ms, alg_bytes, flops):
        gbs = alg_bytes / ms * 1e-6
        print('| %s | %s | %d | %s | %.4f | %.1f | %.0f | %.3f | %.1f |' %
              (op, layer, B, path, ms, alg_bytes / 1e6, gbs, gbs / peak, flops / ms * 1e-9))

    # (name, level, restricted to kept rows?, Cin, Cout, act)
    if args.body:
        convs = [('en0 3->32 @6890 (kept rows)', 0, True, 3, 32), ('en1 32->32 @1723 (kept rows)', 1, True, 32, 32),
                 ('en2 32->64 @431 (kept rows)', 2, True, 32, 64), ('de1 64->64 @431', 2, False, 64, 64),
                 ('de2 64->32 @1723', 1, False, 64, 32), ('de3 32->32 @6890', 0, False, 32, 32),
                 ('de4 32->3 @6890', 0, False, 32, 3)]
        pool_levels = [(2, 64), (1, 64), (0, 32)]
    else:
        pool_levels = [(3, 64), (2, 64), (1, 32), (0, 32)]
    convs = convs if args.body else [('en0 3->32 @17039 (kept rows)', 0, True, 3, 32), ('en1 32->32 @4260 (kept rows)', 1, True, 32, 32),
             ('en2 32->32 @1065 (kept rows)', 2, True, 32, 32), ('en3 32->64 @267 (kept rows)', 3, True, 32, 64),
             ('de1 64->64 @267', 3, False, 64, 64), ('de2 64->32 @1065', 2, False, 64, 32),
             ('de3 32->32 @4260', 1, False, 32, 32), ('de4 32->32 @17039', 0, False, 32, 32),
             ('de5 32->3 @17039', 0, False, 32, 3)]
    for B in [int(b) for b in args.batches.split(',')]:
        for name, lvl, restricted, cin, cout in ([] if args.only == 'pool' else [c for c in convs if args.layer in c[0]]):
            full = spiral_table(sp[lvl])
            tab = restricted_spiral_table(sp[lvl], pool_table(dn[lvl])) if restricted else full
            Vin, R, S = V[lvl], tab.n_rows, tab.seq
            alg = 4.0 * B * (Vin * cin + R * cout)
            flops = 2.0 * B * R * S * cin * cout
            ns = nsets(int(alg))
            xs = [f(B, Vin, cin) for _ in range(ns)]
            ys = [f(B, R, cout) for _ in range(ns)]
            w = f(cout, S * cin) * 0.05
            b = f(cout) * 0.1
            act = cabi.ACT_NONE if cout == 3 else cabi.ACT_ELU
            # tile-staged tcgen05 kernels (csrc/spiral_conv_tile*.cuh) where the table has a tile plan (32 -> 32, patch order)
            tf = tab.tile_fwd() if (cin == 32 and cout == 32) else None
            if tf is not None and cabi.tile_supported(S, 32, 32, tf.rcap, 0):
                wimg_k = torch.empty(cabi.tc_wimg_floats(S, 32, 32), device=DEV)
                cabi.tc_pack_weights(w, wimg_k, S, 32, 32, False, kperm=True)
                ms = timeit(lambda k: cabi.spiralconv_fwd_tile(xs[k], tf, wimg_k, b, ys[k], B, Vin, R, S, 32, 32, act), ns, args.iters)
                row('conv fwd', name, B, 'tcgen05 3xTF32, tile-staged', ms, alg, flops)
                wsk = torch.empty(cabi.spiralconv_bwd_w_workspace(B * R, S, 32, 32) // 4 + 4, device=DEV)
                dWk, dbk = torch.empty(32, S * 32, device=DEV), torch.empty(32, device=DEV)
                ms = timeit(lambda k: cabi.spiralconv_bwd_w_tile(xs[k], tf, ys[k], dWk, dbk, wsk, B, Vin, R, S, 32, 32), ns, args.iters)
                row('conv dW', name, B, 'tcgen05 3xTF32, tile-staged', ms, alg, flops)
            tb_ = tab.tile_bwd() if (cin == 32 and cout == 32) else None
            if tb_ is not None and cabi.tile_supported(S, 32, 32, tb_.rcap, tb_.ecap):
                wimg_kt = torch.empty(cabi.tc_wimg_floats(S, 32, 32), device=DEV)
                cabi.tc_pack_weights(w, wimg_kt, S, 32, 32, True, kperm=True)
                ms = timeit(lambda k: cabi.spiralconv_bwd_x_tile(ys[k], tb_, wimg_kt, None, xs[k], B, R, Vin, S, 32, 32), ns, args.iters)
                row('conv dx', name, B, 'tcgen05 3xTF32, tile-staged', ms, alg, flops)
            # output layer (32 -> 3) on tcgen05: project-then-gather forward, gather-then-project fused backward
            if cin == 32 and cout == 3 and not restricted:
                tpf, tpb = tab.tile_fwd(), tab.tile_bwd()
                if tpf is not None and cabi.narrow_out_fwd_tc_supported(S, cin, cout, tpf.rcap):
                    ms = timeit(lambda k: cabi.narrow_out_fwd_tc(xs[k], tpf, w, b, ys[k], B, Vin, R, S, cin, cout), ns, args.iters)
                    row('conv fwd', name, B, 'tcgen05 3xTF32, project-then-gather', ms, alg, flops)
                if tpb is not None and cabi.narrow_out_bwd_tc_supported(S, cin, cout, tpb.rcap, tpb.ecap):
                    wsq = torch.empty(cabi.narrow_out_bwd_tc_workspace(S, cout) // 4, device=DEV)
                    dWq, dbq = torch.empty(cout, S * cin, device=DEV), torch.empty(cout, device=DEV)
                    dxq = torch.empty(B, Vin, cin, device=DEV)
                    ms = timeit(lambda k: cabi.narrow_out_bwd_tc(ys[k], xs[k], tpb, w, dxq, dWq, dbq, wsq, B, R, Vin, S, cin, cout, True), ns, args.iters)
                    row('conv dx+dW+db', name, B, 'tcgen05 3xTF32, fused gather-then-project', ms, alg + 4.0 * B * Vin * cin, 2 * flops)
                    del dxq
            # forward
            plan = tab.plan_fwd()
            if cin == 32 and cout == 3 and cabi.narrow_out_fwd_supported(S, cin, cout, tab.stage_plan().ucap):
                sp_ = tab.stage_plan()
                ms = timeit(lambda k: cabi.narrow_out_fwd(xs[k], sp_, w, b, ys[k], B, Vin, R, S, cin, cout), ns, args.iters)
                row('conv fwd', name, B, 'fp32 FMA, tile source rows staged in smem', ms, alg, flops)
            if cin in (32, 64) and cabi.tc_supported(S, cin, cout, plan.rcap):
                wimg = torch.empty(cabi.tc_wimg_floats(S, cin, cout), device=DEV)
                cabi.tc_pack_weights(w, wimg, S, cin, cout, False)
                ms = timeit(lambda k: cabi.spiralconv_fwd_tc(xs[k], plan, wimg, b, ys[k], B, Vin, R, S, cin, cout, act), ns, args.iters)
                row('conv fwd', name, B, 'tcgen05 3xTF32', ms, alg, flops)
            elif cabi.narrow_in_supported(Vin, S, cin, cout):
                ms = timeit(lambda k: cabi.narrow_in_fwd(xs[k], tab.idx, w, b, ys[k], B, Vin, R, S, cin, cout, act), ns, args.iters)
                row('conv fwd', name, B, 'fp32 FMA, mesh input resident in smem', ms, alg, flops)
            elif S * cin <= 32 and cout == 32:
                P = torch.empty(B, R, 32, device=DEV)
                Wd = torch.empty(1024, device=DEV)
                wimg = torch.empty(cabi.tc_wimg_floats(1, 32, 32), device=DEV)
                cabi.slot_weight(w, Wd, 0, cout, S, cin)
                cabi.tc_pack_weights(Wd, wimg, 1, 32, 32, False)
                ip = identity_plan(R, DEV)

                def fwd_slot(k):
                    cabi.slot_pack(xs[k], None, tab.idx, P, B, Vin, R, S, cin)
                    cabi.dense_tc(P, ip, wimg, b, None, ys[k], B, R, act)
                ms = timeit(fwd_slot, ns, args.iters)
                row('conv fwd', name, B, 'slot-pack + dense tcgen05', ms, alg, flops)
            elif cin in (32, 64) and cout == 64 and cabi.tc_supported(S, cin, 32, plan.rcap):
                # the engine's path for 64 output channels whose weight image does not fit beside the rings:
                # two passes of 32 output channels, each writing its columns of y (functional._tc_parts)
                wimgs = []
                for n0 in (0, 32):
                    wi = torch.empty(cabi.tc_wimg_floats(S, cin, 32), device=DEV)
                    cabi.tc_pack_weights(w, wi, S, cin, cout, False, n0, 32)
                    wimgs.append(wi)

                def fwd_two(k):
                    for i, n0 in enumerate((0, 32)):
                        cabi.spiralconv_fwd_tc(xs[k], plan, wimgs[i], b[n0:], ys[k].view(-1)[n0:], B, Vin, R, S, cin,
                                               32, act, cout)
                ms = timeit(fwd_two, ns, args.iters)
                row('conv fwd', name, B, 'tcgen05 3xTF32, two 32-channel passes', ms, alg, flops)
            else:
                ms = timeit(lambda k: cabi.spiralconv_fwd(xs[k], tab.idx, w, b, ys[k], B, Vin, R, S, cin, cout, act), ns, args.iters)
                row('conv fwd', name, B, 'fp32 FMA', ms, alg, flops)
            # weight gradient (reads x and dy, writes dW)
            # (the slot-packed paths of the 3-channel layers run a dense 32 x 32 contraction over R or Vin rows: size for both)
            ws = torch.empty(max(cabi.spiralconv_bwd_w_workspace(B * R, S, cin, cout),
                                 cabi.spiralconv_bwd_w_workspace(B * max(R, Vin), 1, 32, 32)) // 4 + 4, device=DEV)
            dW, db = torch.empty(cout, S * cin, device=DEV), torch.empty(cout, device=DEV)
            dWd, dbd = torch.empty(32, 32, device=DEV), torch.empty(32, device=DEV)
            if cabi.narrow_in_supported(Vin, S, cin, cout):
                nws_in = torch.empty(cabi.narrow_in_bwd_w_workspace(S, cin) // 4, device=DEV)
                ms = timeit(lambda k: cabi.narrow_in_bwd_w(xs[k], tab.idx, ys[k], dW, db, nws_in, B, Vin, R, S, cin, cout), ns, args.iters)
                row('conv dW', name, B, 'fp32 FMA, mesh input resident in smem', ms, alg, flops)
            elif S * cin <= 32 and cout == 32:
                # the engine's path: dW = dy^T P with the slot-packed input P of the forward pass
                ipR = identity_plan(R, DEV)
                Pk = torch.randn(B, R, 32, device=DEV)

                def dw_slot_in(k):
                    cabi.spiralconv_bwd_w_tc(Pk, ipR, ys[k], dWd, dbd, ws, B, R, R, 1, 32, 32)
                    cabi.slot_grad(dWd, dbd, dW, db, 0, cout, S, cin)
                ms = timeit(dw_slot_in, ns, args.iters)
                row('conv dW', name, B, 'dense tcgen05 on slot-packed P', ms, alg, flops)
            elif S * cout <= 32 and cin == 32:
                # the engine's path: dW[n, s*32+c] = (G^T x)[s*3+n, c], G shared with the input-gradient pass
                ipV = identity_plan(Vin, DEV)
                Gk = torch.randn(B, Vin, 32, device=DEV)

                def dw_slot_out(k):
                    cabi.spiralconv_bwd_w_tc(xs[k], ipV, Gk, dWd, dbd, ws, B, Vin, Vin, 1, 32, 32)
                    cabi.slot_grad(dWd, dbd, dW, db, 1, cout, S, cout)
                ms = timeit(dw_slot_out, ns, args.iters)
                row('conv dW', name, B, 'dense tcgen05 on slot-packed G (G from the dx pass)', ms, alg, flops)
            elif cin in (32, 64) and cabi.tc_bwd_w_supported(S, cin, cout, plan.rcap):   # 64: 32-channel passes inside
                ms = timeit(lambda k: cabi.spiralconv_bwd_w_tc(xs[k], plan, ys[k], dW, db, ws, B, Vin, R, S, cin, cout), ns, args.iters)
                row('conv dW', name, B, 'tcgen05 3xTF32', ms, alg, flops)
            else:
                ms = timeit(lambda k: cabi.spiralconv_bwd_w(xs[k], tab.idx, ys[k], dW, db, ws, B, Vin, R, S, cin, cout), ns, args.iters)
                row('conv dW', name, B, 'fp32 FMA', ms, alg, flops)
            # backward to input (reads dy, writes dx)
            if cin == 3:
                continue                                   # the first layer has no input gradient
            pb = tab.plan_bwd()
            if cabi.tc_supported(S, cout, cin, pb.rcap):
                wimg_t = torch.empty(cabi.tc_wimg_floats(S, cout, cin), device=DEV)
                cabi.tc_pack_weights(w, wimg_t, S, cin, cout, True)
                ms = timeit(lambda k: cabi.spiralconv_bwd_x_tc(ys[k], pb, wimg_t, None, xs[k], B, R, Vin, S, cout, cin), ns, args.iters)
                row('conv dx', name, B, 'tcgen05 3xTF32', ms, alg, flops)
            elif cout in (32, 64) and cin == 64 and cabi.tc_supported(S, cout, 32, pb.rcap):
                # 64 input channels: two passes of 32, each writing its columns of dx (functional._tc_parts)
                wts = []
                for n0 in (0, 32):
                    wi = torch.empty(cabi.tc_wimg_floats(S, cout, 32), device=DEV)
                    cabi.tc_pack_weights(w, wi, S, cin, cout, True, n0, 32)
                    wts.append(wi)

                def dx_two(k):
                    for i, n0 in enumerate((0, 32)):
                        cabi.spiralconv_bwd_x_tc(ys[k], pb, wts[i], None, xs[k].view(-1)[n0:], B, R, Vin, S, cout, 32, cin)
                ms = timeit(dx_two, ns, args.iters)
                row('conv dx', name, B, 'tcgen05 3xTF32, two 32-channel passes', ms, alg, flops)
            elif S * cout <= 32 and cin == 32:
                G = torch.empty(B, Vin, 32, device=DEV)
                Wd = torch.empty(1024, device=DEV)
                wimg = torch.empty(cabi.tc_wimg_floats(1, 32, 32), device=DEV)
                cabi.slot_weight(w, Wd, 1, cout, S, cout)
                cabi.tc_pack_weights(Wd, wimg, 1, 32, 32, False)
                ip = identity_plan(Vin, DEV)
                cp, cs = tab.inverse()

                def bwd_slot(k):
                    cabi.slot_pack(ys[k], cp, cs, G, B, R, Vin, S, cout)
                    cabi.dense_tc(G, ip, wimg, None, None, xs[k], B, Vin, cabi.ACT_NONE)
                ms = timeit(bwd_slot, ns, args.iters)
                row('conv dx', name, B, 'slot-pack + dense tcgen05', ms, alg, flops)
                del G
                if cabi.narrow_out_bwd_supported(R, S, cin, cout):
                    nws = torch.empty(cabi.narrow_out_bwd_workspace(S, cout) // 4, device=DEV)
                    dxo = torch.empty(B, Vin, cin, device=DEV)
                    cpk = tab.inverse_packed()
                    ms = timeit(lambda k: cabi.narrow_out_bwd(ys[k], xs[k], cp, cs, cpk, w, dxo, dW, db, nws, B, R, Vin, S, cin, cout, True), ns, args.iters)
                    # reads dy and x, writes dx: one more activation-sized tensor than the dx pass alone
                    row('conv dx+dW+db', name, B, 'fused fp32 FMA, dy resident in smem, G in registers', ms, alg + 4.0 * B * Vin * cin, 2 * flops)
                    del dxo
            else:
                wt = torch.empty(cin, S * cout, device=DEV)
                cabi.weight_transpose(w, wt, cout, cin, S)
                cp, cs = tab.inverse()
                ms = timeit(lambda k: cabi.spiralconv_bwd_x(ys[k], cp, cs, wt, None, xs[k], B, R, Vin, S, cout, cin), ns, args.iters)
                row('conv dx', name, B, 'fp32 FMA', ms, alg, flops)
            del xs, ys
        # pools: up-sampling (3 nnz / row) forward and backward; the down-sampling selections are fused
        # into the encoder convolutions (computed at the kept rows only) and have no launch of their own
        for lvl, C in ([] if args.only == 'conv' else pool_levels):
            pt = pool_table(up[lvl])
            Vf, Vc = V[lvl], V[lvl + 1]
            alg = 4.0 * B * C * (Vc + Vf)
            flops = 2.0 * B * 3 * Vf * C
            ns = nsets(int(alg))
            xc = [f(B, Vc, C) for _ in range(ns)]
            xf = [f(B, Vf, C) for _ in range(ns)]
            ms = timeit(lambda k: cabi.pool_ell_fwd(xc[k], pt.ell_col, pt.ell_val, xf[k], B, Vc, Vf, pt.width, C), ns, args.iters)
            row('pool up fwd', '%d->%d C%d' % (Vc, Vf, C), B, 'ELL gather (L2)', ms, alg, flops)
            plan = pt.stage_plan()
            if plan is not None and cabi.pool_stage_supported(C, pt.width, plan.ucap):
                ms = timeit(lambda k: cabi.pool_ell_fwd_staged(xc[k], plan, xf[k], B, Vc, Vf, pt.width, C), ns, args.iters)
                row('pool up fwd', '%d->%d C%d' % (Vc, Vf, C), B, 'ELL, tile source rows staged in smem', ms, alg, flops)
            ms = timeit(lambda k: cabi.csr_rowsum(xf[k], pt.t_ptr, pt.t_row, pt.t_val, None, xc[k], B, Vf, Vc, C), ns, args.iters)
            row('pool up bwd', '%d->%d C%d' % (Vf, Vc, C), B, 'transposed CSR, row per warp', ms, alg, flops)
            del xc, xf
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
