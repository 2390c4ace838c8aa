// Narrow-OUTPUT spiral convolution BACKWARD (the 32 -> 3 output layer; autograd of model.py:34,40 for that layer) on
// tcgen05, one fused pass per tile of 128 input vertices u (GATHER-THEN-PROJECT, the mirror of spiral_conv_tile_out.cuh):
//
//   G[u, s*NO + n] = sum_{v in cell(u, s)} dy[v, n]                         (deterministic in-order scatter-add, inverse table)
//   dx[u, c]       = elu'(x[u, c]) * sum_j G[u, j] * W[j % NO, (j / NO)*32 + c]        (M = u, K = j <= 32, N = c)
//   dW[n, s*32+c]  = sum_u G[u, s*NO + n] * x[u, c]                         (M = c, K = u, N = j: accumulated over the CTA's tiles)
//   db[n]          = sum_u G[u, n]                                          (every dy row lands in exactly one cell of slot 0)
//
// G (27 values per vertex) never leaves the SM: the builder thread of vertex u sums its nine cells from the tile's
// staged dy rows (12 bytes each), writes the hi/lo-split row to TMEM (A operand of the dx contraction) and to a
// shared-memory B operand (MN-major, the layout bw_umma_kernel stages `g` in) for the weight gradient, whose A operand
// is the transposed x tile (thread = channel reads one word of 32 consecutive tile rows: whole-row broadcasts).
// This replaces narrow_out_bwd_kernel (narrow_conv.cuh: fp32 FMA, 1728 FMAs per vertex, 2.75 ms at 1024 meshes).
// Precision: error-compensated 3xTF32 (three N = 32 MMAs per k-step), fp32 accumulation; the weight-gradient
// accumulator is drained every `flush` tiles into registers of the drain warp (the tensor core adds with truncation).
//
// Tile plan: the INVERSE tile plan of the layer's table (tables.tile_plan of inverse_cells: plan_cnt / plan_src = the
// distinct dy rows of a tile, plan_cell / plan_ext = the cells), rcap <= 288.
// Warp roles (24 warps): 0..3 / 4..7 G builders of even / odd tiles (thread = vertex; each set owns one G slot) | 8..11
// and 20..23 dx epilogue of even / odd tiles | 12, 17, 18, 19 x^T transposers, one per TMEM lane quarter, which
// also drain their quarter of the weight-gradient accumulators | 13..15 loaders | 16 TMEM allocation + MMA issue.
// The weight gradient's A operand x^T has only 32 rows (channels) but an MMA spans 128 TMEM lanes: chunk r4 (tile rows
// 32 r4 ..) lives in lane quarter r4 of ITS OWN column range, whose other three quarters are zeroed once -- so four
// warps (one per quarter) share the transposition, accumulator rows 32 q + c collect the chunks r4 = q, and the four
// quarters are added at the end.
// TMEM columns: [0, 64) dx accumulators (2 x 32) | [64, 128) dW accumulators (2 x 32) | [128, 256) G stages (2 x (32 hi +
// 32 lo)) | [256, 512) x^T stages (chunk r4 at 256 + 64 r4: 32 hi + 32 lo).
#pragma once
#include "spiral_conv_tile_out.cuh"
#include "spiral_conv_tile_bw.cuh"

namespace sdvae {
namespace tile {

constexpr int kQThreads = 768;
constexpr int kQEpiWarp0 = 8;
constexpr int kQLoadWarp0 = 13;
constexpr int kQLoadWarps = 3;
constexpr int kQMmaWarp = 16;
constexpr int kQMaxStages = 4;                                        // 4 or 2: two loader warps, warp w owns the stages = w (mod 2)
constexpr int kQXBytes = 128 * 128;                                   // the tile's own x rows

struct OutBwArgs {
    const float* dy;              // [B, rows_v, NO]   gradient w.r.t. the layer's output
    const float* x;               // [B, rows_u, 32]   the layer's input (= ELU output of the previous block)
    const int* plan_cnt;          // inverse tile plan: [L]
    const int* plan_src;          //   [L, rcap/2]
    const uint32_t* plan_cell;    //   [L, S*128]
    const uint16_t* plan_ext;     //   [L, ecap]
    const float* W;               // [NO, S*32]
    float* dx;                    // [B, rows_u, 32]
    float* part;                  // [grid, NO * S * 32] per-CTA partial dW
    float* part_b;                // [grid, NO]          per-CTA partial db
    int B, rows_v, rows_u, L, S, NO, rcap, ecap, nts, flush, gate;
};

struct OutBwCfg {
    static size_t stage_bytes(int S, int rcap, int ecap) { return (size_t)rcap * 16 + kQXBytes + (size_t)S * 512 + (size_t)ecap * 2; }
    static size_t fixed_bytes() { return 1024 + kTBChunk + 2 * (size_t)umma::kGStage + 2 * (size_t)umma::kOutStageBytes + 512 + 4608; }
    static int stages(int S, int rcap, int ecap) {
        const long long budget = 227LL * 1024 - (long long)fixed_bytes();
        long long st = budget / (long long)stage_bytes(S, rcap, ecap);
        return st >= 4 ? 4 : (st >= 2 ? 2 : 0);
    }
    static size_t smem_bytes(int S, int rcap, int ecap, int nts) { return fixed_bytes() + (size_t)nts * stage_bytes(S, rcap, ecap); }
};

#ifndef SDVAE_ABL_QT
#define SDVAE_ABL_QT 0             // compile-time ablation mask of tuning builds: 1 no cell gather, 2 no G_s stores, 4 no dW MMAs,
#endif                             // 8 no transposition, 16 no epilogue work, 32 no G TMEM stores

template <int ST, int NOT>
__global__ void __launch_bounds__(kQThreads, 1)
qt_kernel(const OutBwArgs a) {
    const int S = ST > 0 ? ST : a.S, NO = NOT > 0 ? NOT : a.NO, J = S * NO;
    const int NTS = a.nts;
    const int DY_BYTES = a.rcap * 16;
    const int CELL_OFF = DY_BYTES + kQXBytes;
    const int EXT_OFF = CELL_OFF + S * 512;
    const int STAGE_BYTES = EXT_OFF + a.ecap * 2;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* B_s = smem;                                    // [64][128 B] dx weight image: row c (hi) / 32 + c (lo), K = j
    uint8_t* G_s = B_s + kTBChunk;                          // [2] G tiles as MN-major B operand (hi image | lo image per K atom)
    uint8_t* O_s0 = G_s + 2 * umma::kGStage;                // [2][128][144 B] dx staging, one per epilogue group
    uint8_t* T_s = O_s0 + 2 * umma::kOutStageBytes;             // [NTS] tile stages: dy rows (16 B) | x tile | cell words | ext
    uint64_t* bars = reinterpret_cast<uint64_t*>(T_s + (size_t)NTS * STAGE_BYTES);
    uint64_t* tile_full = bars;                             // [4] loader lanes (async) -> builders, transposers, epilogue
    uint64_t* tile_empty = tile_full + kQMaxStages;         // [4] 4 builder + 4 transposer + 4 epilogue warps -> loader
    uint64_t* adx_full = tile_empty + kQMaxStages;          // [2] builders (4 warps) -> MMA: G in TMEM and in G_s
    uint64_t* adx_empty = adx_full + 2;                     // [2] MMA (commit after the tile's LAST MMA) -> builders
    uint64_t* at_full = adx_empty + 2;                      // [4] transposer -> MMA
    uint64_t* at_empty = at_full + 4;                       // [4] MMA (commit) -> transposer
    uint64_t* tdx_full = at_empty + 4;                      // [2] MMA (commit) -> epilogue
    uint64_t* tdx_empty = tdx_full + 2;                     // [2] epilogue (4 warps) -> MMA
    uint64_t* done_bar = tdx_empty + 2;                     // [2] MMA (commit) -> drain
    uint64_t* drained_bar = done_bar + 2;                   // [2] drain (4 warps) -> MMA
    uint64_t* init_bar = drained_bar + 2;                   // [1] transposers: the foreign quarters of the x^T stages are zero
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(init_bar + 1);
    float* db_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [384][3] per-builder db partials

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kQMaxStages; ++i) { mbar_init(tile_full + i, 32); mbar_init(tile_empty + i, 12); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(adx_full + i, 4); mbar_init(adx_empty + i, 1);
            mbar_init(tdx_full + i, 1); mbar_init(tdx_empty + i, 4);
            mbar_init(done_bar + i, 1); mbar_init(drained_bar + i, 4);
        }
        mbar_init(init_bar, 4);
        for (int i = 0; i < 4; ++i) { mbar_init(at_full + i, 1); mbar_init(at_empty + i, 1); }
        fence_barrier_init();
    }
    if (warp == kQMmaWarp) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    // dx weight image: B[n = c][k = j] = W[j % NO, (j / NO)*32 + c], split hi / lo, K-major 128B-swizzled rows
    for (int t = tid; t < 64 * 32; t += kQThreads) {
        const int j = t & 31, row = t >> 5, c = row & 31, part = row >> 5;
        float w = 0.f;
        if (j < J) w = __ldg(a.W + (size_t)(j % NO) * S * 32 + (j / NO) * 32 + c);
        float hi, lo;
        split_tf32f(w, hi, lo);
        *reinterpret_cast<float*>(B_s + sw128_off(row, j >> 2) + (j & 3) * 4) = part ? lo : hi;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = a.B * a.L;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;
    const int T = a.flush;

    if (warp == kQMmaWarp) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t IDESC_DX = idesc_tf32(kBM, kTNT);
            constexpr uint32_t IDESC_DW = umma::idesc_tf32_bmn(kBM, kTNT);
            const uint64_t wd_hi = smem_desc_sw128(smem_u32(B_s));
            const uint64_t wd_lo = wd_hi + (uint64_t)((32 * 128) >> 4);
            const uint64_t g_desc = umma::smem_desc_mn_sw128(smem_u32(G_s));
            int tf = 0, nfl = 0;
            if (my_tiles > 0) { mbar_wait(init_bar, 0u); tc_fence_after(); }
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                const int sl = it & 1;
                const uint32_t ph = (uint32_t)((it >> 1) & 1);
                const int ab = nfl & 1;
                mbar_wait(tdx_empty + sl, ph ^ 1);
                mbar_wait(adx_full + sl, ph);
                tc_fence_after();
                {   // dx = G W'
                    const uint32_t d_tmem = tmem_base + (uint32_t)(sl * 32);
                    const uint32_t a_hi = tmem_base + (uint32_t)(128 + sl * 64), a_lo = a_hi + 32;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_tf32_ts(d_tmem, a_hi + k * 8, wd_hi + (uint64_t)(2 * k), IDESC_DX, k != 0);
                        umma_tf32_ts(d_tmem, a_hi + k * 8, wd_lo + (uint64_t)(2 * k), IDESC_DX, 1u);
                        umma_tf32_ts(d_tmem, a_lo + k * 8, wd_hi + (uint64_t)(2 * k), IDESC_DX, 1u);
                    }
                    umma_commit(tdx_full + sl);
                }
                // dW += x^T G, four chunks of 32 tile rows
                if (tf == 0 && nfl >= 2) {
                    mbar_wait(drained_bar + ab, (uint32_t)(((nfl >> 1) - 1) & 1));
                    tc_fence_after();
                }
                const bool last_of_group = (tf == T - 1) || (it == my_tiles - 1);
                const uint32_t dw_tmem = tmem_base + (uint32_t)(64 + ab * 32);
#pragma unroll 1
                for (int r4 = 0; r4 < 4; ++r4) {
                    mbar_wait(at_full + r4, (uint32_t)(it & 1));
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + (uint32_t)(256 + r4 * 64);
                    const uint64_t bd0 = g_desc + (uint64_t)((sl * umma::kGStage + r4 * 4 * 2048) >> 4);
                    if (!(SDVAE_ABL_QT & 4)) {
                        umma_bw_2k(dw_tmem, a_hi, bd0, IDESC_DW, (uint32_t)(tf | r4));
                        umma_bw_2k(dw_tmem, a_hi + 16, bd0 + 256, IDESC_DW, 1u);
                    }
                    umma_commit(at_empty + r4);
                }
                umma_commit(adx_empty + sl);               // G stage (TMEM) and G_s[sl] are free again
                if (last_of_group) { umma_commit(done_bar + ab); tf = 0; ++nfl; } else ++tf;
            }
        }
        __syncwarp();
    } else if (warp >= kQLoadWarp0 && warp < kQLoadWarp0 + kQLoadWarps) {
        // ================= loaders: warp lw < 2 takes the tiles it = lw (mod 2) into stage it % NTS (NTS even): one producer
        // per stage barrier (a parity wait cannot tell phases two apart); asynchronous arrival, two tiles in flight per warp ====
        const int lw = warp - kQLoadWarp0;
        long long t0 = (long long)blockIdx.x + (long long)lw * gridDim.x;
        int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
        const int n_tail16 = (S * 512 + a.ecap * 2) >> 4;
#pragma unroll 1
        for (int it = lw; lw < 2 && it < my_tiles; it += 2) {
            const int ts = it % NTS;
            const uint32_t stage_a = smem_u32(T_s) + (uint32_t)ts * (uint32_t)STAGE_BYTES;
            // the tile's source list first (independent 16-byte loads; a dependent index load per row made the loader
            // the slowest role): lane (rsub = lane >> 3, q = lane & 7) owns the staged dy rows e = 32 j + 4 q + rsub
            constexpr int PV = kTMaxRcap / 32;
            PlanRegs<PV> now;
            plan_fetch(now, a.plan_cnt, a.plan_src, jt, 1, 0, a.rcap, lane >> 3);
            const char* dyb = reinterpret_cast<const char*>(a.dy + (size_t)b * a.rows_v * NO);
            const char* xb = reinterpret_cast<const char*>(a.x + ((size_t)b * a.rows_u + (size_t)jt * kBM) * 32);
            const int nvalid = min(kBM, a.rows_u - jt * kBM);
            const char* cell_g = reinterpret_cast<const char*>(a.plan_cell + (size_t)jt * S * 128);
            const char* ext_g = reinterpret_cast<const char*>(a.plan_ext + (size_t)jt * a.ecap);
            mbar_wait_relaxed(tile_empty + ts, (uint32_t)(((it / NTS) & 1) ^ 1));
#pragma unroll
            for (int j = 0; j < PV; ++j) {
                const int e = 32 * j + 4 * (lane & 7) + (lane >> 3);
                if (e < now.n) {
                    const uint32_t w[4] = {now.w[j].x, now.w[j].y, now.w[j].z, now.w[j].w};
                    const int t = lane & 7;
                    uint32_t ww = w[0];
                    ww = (t >> 1) == 1 ? w[1] : ww; ww = (t >> 1) == 2 ? w[2] : ww; ww = (t >> 1) == 3 ? w[3] : ww;
                    const uint32_t v = (t & 1) ? (ww >> 16) : (ww & 0xffffu);
                    const char* srow = dyb + (size_t)v * NO * 4;
                    const uint32_t dst = stage_a + (uint32_t)e * 16u;
                    for (int n = 0; n < NO; ++n)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + 4u * n), "l"(srow + 4 * n));
                }
            }
            // the tile's own x rows (rows past the mesh shadow its last row)
#pragma unroll 4
            for (int i = lane; i < kBM * 8; i += 32) {
                const int row = i >> 3, q = i & 7;
                const int rr = row < nvalid ? row : nvalid - 1;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                             ::"r"(stage_a + (uint32_t)DY_BYTES + (uint32_t)i * 16u), "l"(xb + (size_t)rr * 128 + q * 16));
            }
#pragma unroll 1
            for (int i = lane; i < n_tail16; i += 32) {
                const int off = i * 16;
                const char* src = off < S * 512 ? cell_g + off : ext_g + (off - S * 512);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(stage_a + (uint32_t)CELL_OFF + (uint32_t)off), "l"(src));
            }
            cp_async_arrive_noinc(smem_u32(tile_full + ts));
            for (int k = 0; k < 2; ++k) {
                b += db; jt += djt;
                if (jt >= a.L) { jt -= a.L; ++b; }
            }
        }
    } else if (warp < 8) {
        // ================= G builders: thread = vertex u of the tile.  Two sets, each the ONLY producer of its G slot
        // (TMEM stage + G_s buffer): a third set sharing the slots would be two barrier phases away from the slot's
        // previous producer, which a parity wait cannot tell apart =================
        const int set = warp >> 2;                             // tiles it = set (mod 2), slot = set
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t cell_idx = (uint32_t)((r >> 5) * 32 + (r & 7) * 4 + ((r >> 3) & 3)) * 4u;
        const uint32_t t_g0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + 128u;
        float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
        long long t0 = (long long)blockIdx.x + (long long)set * gridDim.x;
        int jt = (int)(t0 % a.L);
        const int djt2 = (2 * djt) % a.L;
#pragma unroll 1
        for (int it = set; it < my_tiles; it += 2) {
            const int sl = set;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            const uint32_t t_g = t_g0 + (uint32_t)(sl * 64);
            uint8_t* gs = G_s + (size_t)sl * umma::kGStage;
            const bool valid = jt * kBM + r < a.rows_u;
            jt += djt2; if (jt >= a.L) jt -= a.L;
            const int ts = it % NTS;
            const uint32_t stage_a = smem_u32(T_s) + (uint32_t)ts * (uint32_t)STAGE_BYTES;
            mbar_wait_a<64>(smem_u32(tile_full + ts), (uint32_t)((it / NTS) & 1));
            float g[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = 0.f;
            if (!(SDVAE_ABL_QT & 1)) {
                // all nine cell words, then all nine first rows (independent loads in flight together; an empty cell's word
                // points at staged row 0, its value is dropped), then -- rarely -- the cells' further rows in order
                uint32_t w[9];
#pragma unroll
                for (int s = 0; s < 9; ++s) w[s] = s < S ? lds32u(stage_a + (uint32_t)CELL_OFF + cell_idx + (uint32_t)s * 512u) : 0u;
#pragma unroll
                for (int s = 0; s < 9; ++s) {
                    if (s < S) {
                        const float4 d = lds128(stage_a + (((w[s] & 0xffffu) >> 7) << 4));
                        const bool any = ((w[s] >> 16) & 0x1fu) != 0u;
                        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                        for (int n = 0; n < 4; ++n)
                            if (n < NO) g[s * NO + n] = any ? dv[n] : 0.f;
                    }
                }
#pragma unroll
                for (int s = 0; s < 9; ++s) {
                    if (s < S) {
                        const int cnt = (int)((w[s] >> 16) & 0x1fu);
                        if (cnt > 1) {
                            uint32_t ea = stage_a + (uint32_t)EXT_OFF + ((w[s] >> 21) << 1);
#pragma unroll 1
                            for (int e = 1; e < cnt; ++e, ea += 2) {   // in-order sum: deterministic scatter-add
                                const float4 d2 = lds128(stage_a + ((lds16u(ea) >> 7) << 4));
                                const float dv[4] = {d2.x, d2.y, d2.z, d2.w};
#pragma unroll
                                for (int n = 0; n < 4; ++n)
                                    if (n < NO) g[s * NO + n] += dv[n];
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(smem_u32(tile_empty + ts));   // the builders are done with the stage
#pragma unroll
            for (int n = 0; n < 4; ++n)
                if (n < NO && valid) dbacc[n] += g[n];
            mbar_wait_a<32>(smem_u32(adx_empty + sl), ph ^ 1);
            __syncwarp();
            tc_fence_after();
            float lo[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 v2 = make_float2(g[j], g[j + 1]);
                const float2 hi = make_float2(__uint_as_float(__float_as_uint(v2.x) & 0xffffe000u),
                                              __uint_as_float(__float_as_uint(v2.y) & 0xffffe000u));
                const float2 l2 = sub2(v2, hi);
                g[j] = hi.x; g[j + 1] = hi.y; lo[j] = l2.x; lo[j + 1] = l2.y;
            }
            if (!(SDVAE_ABL_QT & 32)) {
                tmem_st32(t_g, g);
                tmem_st32(t_g + 32, lo);
            }
            if (!(SDVAE_ABL_QT & 2))
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint8_t* dst = gs + umma::g_off(r, q);
                *reinterpret_cast<float4*>(dst) = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
                *reinterpret_cast<float4*>(dst + 1024) = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
            tmem_st_wait();
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(smem_u32(adx_full + sl));
        }
        for (int n = 0; n < 3; ++n) db_s[((set * 4 + q4) * 32 + lane) * 3 + n] = n < NO ? dbacc[n] : 0.f;
    } else if ((warp >= kQEpiWarp0 && warp < kQEpiWarp0 + 4) || warp >= 20) {
        // ================= dx epilogue: TMEM -> staging -> elu' gate -> coalesced store; two groups of four warps, group g
        // owns the accumulator slot g and takes the tiles it = g (mod 2) =================
        const int grp = warp >= 20 ? 1 : 0;
        const int q4 = warp & 3;
        uint8_t* O_s = O_s0 + (size_t)grp * umma::kOutStageBytes;
        long long t0 = (long long)blockIdx.x + (long long)grp * gridDim.x;
        int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
#pragma unroll 1
        for (int it = grp; it < my_tiles; it += 2) {
            const int sl = grp;
            const int ts = it % NTS;
            mbar_wait_relaxed(tdx_full + sl, (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(sl * 32);
            float v0[16], v1[16];
            tmem_ld16(t_row, v0);
            tmem_ld16(t_row + 16, v1);
            tmem_ld_wait();
            tc_fence_before();
            warp_arrive(tdx_empty + sl, lane);
            float4* srow = reinterpret_cast<float4*>(O_s + (q4 * 32 + lane) * umma::kOutRowBytes);
            if (!(SDVAE_ABL_QT & 16))
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                srow[j >> 2] = make_float4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
                srow[4 + (j >> 2)] = make_float4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
            }
            __syncwarp();
            mbar_wait_a<64>(smem_u32(tile_full + ts), (uint32_t)((it / NTS) & 1));   // the x tile (gate) is in the stage
            const uint8_t* xs = T_s + (size_t)ts * STAGE_BYTES + DY_BYTES;
            const int piece = lane & 7;
            if (!(SDVAE_ABL_QT & 16))
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int lr2 = q4 * 32 + 4 * k + (lane >> 3);
                float4 t = *reinterpret_cast<const float4*>(O_s + lr2 * umma::kOutRowBytes + piece * 16);
                if (a.gate) {
                    const float4 gt = *reinterpret_cast<const float4*>(xs + lr2 * 128 + piece * 16);
                    t.x *= elu_grad_from_out(gt.x); t.y *= elu_grad_from_out(gt.y);
                    t.z *= elu_grad_from_out(gt.z); t.w *= elu_grad_from_out(gt.w);
                }
                const int r2 = jt * kBM + lr2;
                if (r2 < a.rows_u)
                    *reinterpret_cast<float4*>(a.dx + ((size_t)b * a.rows_u + r2) * 32 + piece * 4) = t;
            }
            warp_arrive(tile_empty + ts, lane);
            for (int k = 0; k < 2; ++k) {
                b += db; jt += djt;
                if (jt >= a.L) { jt -= a.L; ++b; }
            }
        }
    } else if (warp == 12 || (warp >= 17 && warp <= 19)) {
        // ================= x^T transposers: warp of lane quarter q4 serves chunk r4 = q4 of every tile =================
        const int r4 = warp & 3;
        const uint32_t t_q = tmem_base + ((uint32_t)(r4 * 32) << 16);
        {   // once: zero this quarter's lanes in the three foreign chunk stages
            float z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0.f;
            for (int o = 0; o < 4; ++o)
                if (o != r4) { tmem_st32(t_q + (uint32_t)(256 + o * 64), z); tmem_st32(t_q + (uint32_t)(256 + o * 64 + 32), z); }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(init_bar);
        }
        int jt = (int)blockIdx.x % a.L;
        const uint32_t t_a = t_q + (uint32_t)(256 + r4 * 64);
        // This warp also drains its lane quarter of the weight-gradient accumulators (thread = channel c keeps
        // dWd[c][j] of its quarter in registers): flush group f is drained while the tiles of group f + 1 are in
        // flight -- the MMA thread needs the set back only for group f + 2.
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        const int nflush = (my_tiles + T - 1) / T;
        int fdone = 0;
        auto drain = [&](int f) {
            const int ab = f & 1;
            mbar_wait(done_bar + ab, (uint32_t)((f >> 1) & 1));
            tc_fence_after();
            float v0[16], v1[16];
            tmem_ld16(t_q + (uint32_t)(64 + ab * 32), v0);
            tmem_ld16(t_q + (uint32_t)(64 + ab * 32 + 16), v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(drained_bar + ab);
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc[j] += v0[j]; acc[16 + j] += v1[j]; }
        };
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int ts = it % NTS;
            const int nvalid = min(kBM, a.rows_u - jt * kBM);
            const uint32_t xs = smem_u32(T_s) + (uint32_t)ts * (uint32_t)STAGE_BYTES + (uint32_t)DY_BYTES + (uint32_t)lane * 4u;
            mbar_wait_a<64>(smem_u32(tile_full + ts), (uint32_t)((it / NTS) & 1));
            float v[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float t = (SDVAE_ABL_QT & 8) ? 1.f : lds32(xs + (uint32_t)(32 * r4 + j) * 128u);
                v[j] = 32 * r4 + j < nvalid ? t : 0.f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(smem_u32(tile_empty + ts));
            mbar_wait_a<32>(smem_u32(at_empty + r4), (uint32_t)((it & 1) ^ 1));
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 v2 = make_float2(v[j], v[j + 1]);
                const float2 hi = make_float2(__uint_as_float(__float_as_uint(v2.x) & 0xffffe000u),
                                              __uint_as_float(__float_as_uint(v2.y) & 0xffffe000u));
                const float2 l2 = sub2(v2, hi);
                v[j] = hi.x; v[j + 1] = hi.y; lo[j] = l2.x; lo[j + 1] = l2.y;
            }
            tc_fence_after();
            if (!(SDVAE_ABL_QT & 8)) {
                tmem_st32(t_a, v);
                tmem_st32(t_a + 32, lo);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(smem_u32(at_full + r4));
            jt += djt; if (jt >= a.L) jt -= a.L;
            // groups that ended at least one tile ago (group f covers the tiles f*T .. f*T + T - 1)
            while (fdone < nflush && (fdone + 1) * T - 1 < it) { drain(fdone); ++fdone; }
        }
        while (fdone < nflush) { drain(fdone); ++fdone; }
        // the four quarters are added after the CTA barrier below; G_s is free (every MMA has completed): [q4][j][c]
        float* qs = reinterpret_cast<float*>(G_s) + r4 * 1024;
#pragma unroll
        for (int j = 0; j < 32; ++j) qs[j * 32 + lane] = acc[j];
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 12) {                                          // dW partial of this CTA: quarters in order (deterministic)
        const float* qs = reinterpret_cast<const float*>(G_s);
        float* P = a.part + (size_t)blockIdx.x * NO * S * 32;
        for (int j = 0; j < J; ++j) {
            const float t = ((qs[j * 32 + lane] + qs[1024 + j * 32 + lane]) + qs[2048 + j * 32 + lane]) + qs[3072 + j * 32 + lane];
            P[(size_t)(j % NO) * S * 32 + (j / NO) * 32 + lane] = t;
        }
    }
    if (tid < NO) {                                            // db partial of this CTA: the 256 builder threads in order
        float t = 0.f;
        for (int i = 0; i < 256; ++i) t += db_s[i * 3 + tid];
        a.part_b[(size_t)blockIdx.x * NO + tid] = t;
    }
    if (warp == kQMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace tile
}  // namespace sdvae
