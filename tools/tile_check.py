#!/usr/bin/env python
"""Development check + timing of the tile-staged tcgen05 SpiralConv kernels (csrc/spiral_conv_tile.cuh) against a
float64 evaluation on the GPU and against the per-slot-gather tcgen05 kernels, on the patch-ordered craniofacial
template.   python tools/tile_check.py [--levels 0,1] [--B 1024] [--iters 5] [--skip-check]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdvae_b200 import cabi, fixtures as fx, tables as tb

DEV = 'cuda:0'


def nerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def print_prof(tag):
    if not (int(os.environ.get('SDVAE_DBG', '0')) & 32) and not os.environ.get('SDVAE_PROF'):
        return
    import ctypes
    buf = (ctypes.c_longlong * 64)()
    cabi.load().sdvae_debug_read_prof(buf)
    p = list(buf)
    ch = max(p[3], 1)
    print('  prof %s (CTA 0, clk; per chunk in brackets) chunks %d' % (tag, p[3]))
    print('    mma      total %d [%.0f]  wait a_full %d [%.0f]  wait t_empty %d' % (p[0], p[0] / ch, p[1], p[1] / ch, p[2]))
    print('    loader0  total %d  wait tile_empty %d  wait copies %d' % (p[8], p[9], p[10]))
    print('    epilogue total %d  wait t_full %d' % (p[16], p[17]))
    u = max(ch / 4, 1)
    print('    split    total %d [%.0f per unit]: wait tile_full %.0f  gather+split %.0f  wait::st(prev)+arrive %.0f  wait a_empty %.0f  STTM issue %.0f  advance %.0f'
          % (p[24], p[24] / u, p[25] / u, p[26] / u, p[27] / u, p[28] / u, p[29] / u, p[30] / u), flush=True)


def print_timeline(tag):
    if not (int(os.environ.get('SDVAE_DBG', '0')) & 64):
        return
    import ctypes
    buf = (ctypes.c_longlong * 256)()
    if not hasattr(cabi.load(), 'sdvae_debug_read_timeline'):
        return
    cabi.load().sdvae_debug_read_timeline(buf)
    t = list(buf)
    t0 = min(x for x in t if x > 0)
    print('  timeline %s (CTA 0, clk relative; chunks 400..415 = rounds 200..207)' % tag)
    for r in range(8):
        print('    round %d  mma-hi: a_full seen %6d, 8 MMAs issued %6d | mma-lo: %6d, %6d' % (200 + r, t[r * 4] - t0, t[r * 4 + 1] - t0, t[64 + r * 4] - t0, t[64 + r * 4 + 1] - t0))
    for g in range(16):
        x = t[128 + g * 4:128 + g * 4 + 4]
        if x[0] > 0:
            print('    chunk %d (set %d, q4 0): split done / wait a_empty %6d, a_empty seen %6d, STTM done %6d, a_full arrive %6d' % (400 + g, g % 4, x[0] - t0, x[1] - t0, x[2] - t0, x[3] - t0))


def ev_time(fn, iters, nbuf):
    fn(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--levels', default='0,1')
    ap.add_argument('--B', type=int, default=1024)
    ap.add_argument('--iters', type=int, default=6)
    ap.add_argument('--skip-check', action='store_true')
    ap.add_argument('--skip-old', action='store_true')
    ap.add_argument('--only', default='fwd,dx,dw')
    ap.add_argument('--lib', default='', help='path of a variant build of libsdvae_b200.so (tuning runs)')
    a = ap.parse_args()
    if a.lib:
        cabi.LIB_PATH = os.path.abspath(a.lib)
    tabs = fx.craniofacial_tables().renumbered(128)[0]
    S = 9
    for lvl in [int(t) for t in a.levels.split(',')]:
        idx = tabs.spiral_tensors()[lvl]
        V = idx.shape[0]
        tab = tb.spiral_table(idx.to(DEV))
        t0 = time.time()
        pf, pb = tab.tile_fwd(), tab.tile_bwd()
        print('level %d V=%d: tile plans rcap fwd %d bwd %d ecap %d (%.2fs)' % (lvl, V, pf.rcap, pb.rcap, pb.ecap, time.time() - t0), flush=True)
        assert cabi.tile_supported(S, 32, 32, pf.rcap, 0) and cabi.tile_supported(S, 32, 32, pb.rcap, pb.ecap)
        g = torch.Generator(device='cpu').manual_seed(lvl)
        W = (torch.randn(32, S * 32, generator=g) * 0.1).to(DEV)
        bias = (torch.randn(32, generator=g) * 0.2).to(DEV)
        wimg = torch.empty(cabi.tc_wimg_floats(S, 32, 32), device=DEV)
        wimg_o = torch.empty_like(wimg)
        wimg_t = torch.empty_like(wimg)
        wimg_to = torch.empty_like(wimg)
        cabi.tc_pack_weights(W, wimg, S, 32, 32, False, kperm=True)
        cabi.tc_pack_weights(W, wimg_o, S, 32, 32, False)
        cabi.tc_pack_weights(W, wimg_t, S, 32, 32, True, kperm=True)
        cabi.tc_pack_weights(W, wimg_to, S, 32, 32, True)
        idx_d = idx.to(DEV)
        if not a.skip_check:
            for B in (3, 41):
                x = torch.randn(B, V, 32, generator=g).to(DEV)
                if 'fwd' in a.only:
                    for act in (1, 0):
                        y = torch.full((B, V, 32), float('nan'), device=DEV)
                        cabi.spiralconv_fwd_tile(x, pf, wimg, bias, y, B, V, V, S, 32, 32, act)
                        torch.cuda.synchronize()
                        ref = (x.double()[:, idx_d.view(-1)].view(B, V, S * 32) @ W.double().t()) + bias.double()
                        if act:
                            ref = torch.where(ref > 0, ref, torch.expm1(ref))
                        y2 = torch.full((B, V, 32), float('nan'), device=DEV)
                        cabi.spiralconv_fwd_tile(x, pf, wimg, bias, y2, B, V, V, S, 32, 32, act)
                        torch.cuda.synchronize()
                        print('  fwd  B=%d act=%d: normwise err vs fp64 %.3e   deterministic %s' % (B, act, nerr(y, ref), torch.equal(y, y2)), flush=True)
                        del ref
                if 'dx' in a.only:
                    dpre = torch.randn(B, V, 32, generator=g).to(DEV)
                    gate = torch.randn(B, V, 32, generator=g).to(DEV)
                    for use_gate in (False, True):
                        dx = torch.full((B, V, 32), float('nan'), device=DEV)
                        cabi.spiralconv_bwd_x_tile(dpre, pb, wimg_t, gate if use_gate else None, dx, B, V, V, S, 32, 32)
                        torch.cuda.synchronize()
                        dG = (dpre.double() @ W.double()).view(B, V * S, 32)
                        ref = torch.zeros(B, V, 32, dtype=torch.float64, device=DEV)
                        ref.index_add_(1, idx_d.view(-1), dG)
                        if use_gate:
                            ref = ref * torch.where(gate > 0, torch.ones_like(gate), gate + 1).double()
                        dx2 = torch.full((B, V, 32), float('nan'), device=DEV)
                        cabi.spiralconv_bwd_x_tile(dpre, pb, wimg_t, gate if use_gate else None, dx2, B, V, V, S, 32, 32)
                        torch.cuda.synchronize()
                        print('  dx   B=%d gate=%d: normwise err vs fp64 %.3e   deterministic %s' % (B, use_gate, nerr(dx, ref), torch.equal(dx, dx2)), flush=True)
                        del ref, dG
                if 'dw' in a.only:
                    gg_ = torch.randn(B, V, 32, generator=g).to(DEV)
                    ws = torch.empty(cabi.spiralconv_bwd_w_workspace(B * V, S, 32, 32) // 4 + 4, device=DEV)
                    res = []
                    for _ in range(2):
                        dW = torch.full((32, S * 32), float('nan'), device=DEV); db = torch.full((32,), float('nan'), device=DEV)
                        cabi.spiralconv_bwd_w_tile(x, pf, gg_, dW, db, ws, B, V, V, S, 32, 32)
                        torch.cuda.synchronize()
                        res.append((dW, db))
                    A = x.double()[:, idx_d.view(-1)].view(B * V, S * 32)
                    refW = gg_.double().view(B * V, 32).t() @ A
                    refb = gg_.double().sum((0, 1))
                    print('  dW   B=%d: normwise err vs fp64 dW %.3e db %.3e   deterministic %s' % (B, nerr(res[0][0], refW), nerr(res[0][1], refb), torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])), flush=True)
                    del A
        if 'out' in a.only:
            # narrow-output layer (32 -> 3) by project-then-gather (csrc/spiral_conv_tile_out.cuh)
            assert cabi.narrow_out_fwd_tc_supported(S, 32, 3, pf.rcap), 'pt_kernel: unsupported plan'
            W3 = (torch.randn(3, S * 32, generator=g) * 0.1).to(DEV)
            b3 = (torch.randn(3, generator=g) * 0.2).to(DEV)
            if not a.skip_check:
                for B in (1, 3, 41):
                    x = torch.randn(B, V, 32, generator=g).to(DEV)
                    ys = []
                    for _ in range(2):
                        y = torch.full((B, V, 3), float('nan'), device=DEV)
                        cabi.narrow_out_fwd_tc(x, pf, W3, b3, y, B, V, V, S, 32, 3)
                        torch.cuda.synchronize()
                        ys.append(y)
                    ref = (x.double()[:, idx_d.view(-1)].view(B, V, S * 32) @ W3.double().t()) + b3.double()
                    print('  out  B=%d: normwise err vs fp64 %.3e   deterministic %s' % (B, nerr(ys[0], ref), torch.equal(ys[0], ys[1])), flush=True)
                    del ref
            Bt = a.B
            nb3 = max(2, int(np.ceil(300e6 / (Bt * V * 140))) + 1)
            xs3 = [torch.randn(Bt, V, 32, device=DEV) for _ in range(nb3)]
            ys3 = [torch.empty(Bt, V, 3, device=DEV) for _ in range(nb3)]
            alg3 = Bt * V * 140
            ms = ev_time(lambda i: cabi.narrow_out_fwd_tc(xs3[i], pf, W3, b3, ys3[i], Bt, V, V, S, 32, 3), a.iters, nb3)
            print('  %-34s B=%d  %.4f ms  %.0f GB/s alg (%.3f of 6556)' % ('out project-then-gather (tcgen05)', Bt, ms, alg3 / ms / 1e6, alg3 / ms / 1e6 / 6556.2), flush=True)
            sp_ = tab.stage_plan()
            if cabi.narrow_out_fwd_supported(S, 32, 3, sp_.ucap):
                ms = ev_time(lambda i: cabi.narrow_out_fwd(xs3[i], sp_, W3, b3, ys3[i], Bt, V, V, S, 32, 3), a.iters, nb3)
                print('  %-34s B=%d  %.4f ms  %.0f GB/s alg (%.3f of 6556)' % ('out fp32 FMA, staged rows', Bt, ms, alg3 / ms / 1e6, alg3 / ms / 1e6 / 6556.2), flush=True)
            del xs3, ys3
        if 'bwo' in a.only:
            # narrow-output layer backward (32 -> 3), gather-then-project (csrc/spiral_conv_tile_out_bw.cuh)
            assert cabi.narrow_out_bwd_tc_supported(S, 32, 3, pb.rcap, pb.ecap), 'qt_kernel: unsupported plan'
            W3 = (torch.randn(3, S * 32, generator=g) * 0.1).to(DEV)
            ws3 = torch.empty(cabi.narrow_out_bwd_tc_workspace(S, 3) // 4, device=DEV)
            if not a.skip_check:
                for B in (1, 3, 41):
                    x = torch.randn(B, V, 32, generator=g).to(DEV)
                    dy = torch.randn(B, V, 3, generator=g).to(DEV)
                    res = []
                    for _ in range(2):
                        dx = torch.full((B, V, 32), float('nan'), device=DEV)
                        dW = torch.full((3, S * 32), float('nan'), device=DEV); db_ = torch.full((3,), float('nan'), device=DEV)
                        cabi.narrow_out_bwd_tc(dy, x, pb, W3, dx, dW, db_, ws3, B, V, V, S, 32, 3, True)
                        torch.cuda.synchronize()
                        res.append((dx, dW, db_))
                    dG = (dy.double() @ W3.double()).view(B, V * S, 32)
                    rdx = torch.zeros(B, V, 32, dtype=torch.float64, device=DEV)
                    rdx.index_add_(1, idx_d.view(-1), dG)
                    rdx = rdx * torch.where(x > 0, torch.ones_like(x), x + 1).double()
                    A = x.double()[:, idx_d.view(-1)].view(B * V, S * 32)
                    rdW = dy.double().view(B * V, 3).t() @ A
                    rdb = dy.double().sum((0, 1))
                    det = all(torch.equal(res[0][i], res[1][i]) for i in range(3))
                    print('  bwo  B=%d: normwise err vs fp64 dx %.3e dW %.3e db %.3e   deterministic %s'
                          % (B, nerr(res[0][0], rdx), nerr(res[0][1], rdW), nerr(res[0][2], rdb), det), flush=True)
                    del A, dG, rdx
            Bt = a.B
            nb3 = max(2, int(np.ceil(300e6 / (Bt * V * 268))) + 1)
            xs3 = [torch.randn(Bt, V, 32, device=DEV) for _ in range(nb3)]
            dys3 = [torch.randn(Bt, V, 3, device=DEV) for _ in range(nb3)]
            dxs3 = [torch.empty(Bt, V, 32, device=DEV) for _ in range(nb3)]
            dW = torch.empty(3, S * 32, device=DEV); db_ = torch.empty(3, device=DEV)
            alg3 = Bt * V * 268
            ms = ev_time(lambda i: cabi.narrow_out_bwd_tc(dys3[i], xs3[i], pb, W3, dxs3[i], dW, db_, ws3, Bt, V, V, S, 32, 3, True), a.iters, nb3)
            print('  %-34s B=%d  %.4f ms  %.0f GB/s alg (%.3f of 6556)' % ('out bwd gather-then-project (tc)', Bt, ms, alg3 / ms / 1e6, alg3 / ms / 1e6 / 6556.2), flush=True)
            if cabi.narrow_out_bwd_supported(V, S, 32, 3):
                nws = torch.empty(cabi.narrow_out_bwd_workspace(S, 3) // 4, device=DEV)
                cp, cs = tab.inverse()
                cpk = tab.inverse_packed()
                ms = ev_time(lambda i: cabi.narrow_out_bwd(dys3[i], xs3[i], cp, cs, cpk, W3, dxs3[i], dW, db_, nws, Bt, V, V, S, 32, 3, True), a.iters, nb3)
                print('  %-34s B=%d  %.4f ms  %.0f GB/s alg (%.3f of 6556)' % ('out bwd fused fp32 FMA', Bt, ms, alg3 / ms / 1e6, alg3 / ms / 1e6 / 6556.2), flush=True)
            del xs3, dys3, dxs3
        # ---- timing ----
        B = a.B
        nbuf = max(2, int(np.ceil(300e6 / (B * V * 128))) + 1)          # rotate buffers larger than L2
        xs = [torch.randn(B, V, 32, device=DEV) for _ in range(nbuf)]
        ys = [torch.empty(B, V, 32, device=DEV) for _ in range(nbuf)]
        alg = 2 * B * V * 128
        flops = 2.0 * B * V * 288 * 32
        def rep(name, ms):
            print('  %-34s B=%d  %.4f ms  %.0f GB/s alg (%.3f of 6556)  %.1f TFLOP/s' % (name, B, ms, alg / ms / 1e6, alg / ms / 1e6 / 6556.2, flops / ms / 1e9), flush=True)
        if 'fwd' in a.only:
            rep('fwd tile-staged', ev_time(lambda i: cabi.spiralconv_fwd_tile(xs[i], pf, wimg, bias, ys[i], B, V, V, S, 32, 32, 1), a.iters, nbuf))
            print_prof('fwd'); print_timeline('fwd')
            if not a.skip_old:
                po = tab.plan_fwd()
                rep('fwd per-slot gather (gc_umma)', ev_time(lambda i: cabi.spiralconv_fwd_tc(xs[i], po, wimg_o, bias, ys[i], B, V, V, S, 32, 32, 1), a.iters, nbuf))
        if 'dx' in a.only:
            rep('dx  tile-staged', ev_time(lambda i: cabi.spiralconv_bwd_x_tile(xs[i], pb, wimg_t, None, ys[i], B, V, V, S, 32, 32), a.iters, nbuf))
            print_prof('dx')
            if not a.skip_old:
                po = tab.plan_bwd()
                rep('dx  per-slot gather (gc_umma)', ev_time(lambda i: cabi.spiralconv_bwd_x_tc(xs[i], po, wimg_to, None, ys[i], B, V, V, S, 32, 32), a.iters, nbuf))
        if 'dw' in a.only:
            po = tab.plan_fwd()
            ws = torch.empty(cabi.spiralconv_bwd_w_workspace(B * V, S, 32, 32) // 4 + 4, device=DEV)
            dW = torch.empty(32, S * 32, device=DEV); db = torch.empty(32, device=DEV)
            rep('dW  tile-staged', ev_time(lambda i: cabi.spiralconv_bwd_w_tile(xs[i], pf, ys[(i + 1) % nbuf], dW, db, ws, B, V, V, S, 32, 32), a.iters, nbuf))
            if not a.skip_old:
                rep('dW  per-slot gather (bw_umma)', ev_time(lambda i: cabi.spiralconv_bwd_w_tc(xs[i], po, ys[(i + 1) % nbuf], dW, db, ws, B, V, V, S, 32, 32), a.iters, nbuf))
        del xs, ys


if __name__ == '__main__':
    main()
