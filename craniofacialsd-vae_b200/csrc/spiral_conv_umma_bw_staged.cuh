// EXPERIMENTAL (compiled, NOT dispatched by the engine, the autograd functions or bench.py; reached only through
// sdvae_spiralconv_bwd_w_tc_staged, which tests/ exercise only when SDVAE_EXPERIMENTAL=1): the tcgen05 weight
// gradient of spiral_conv_umma_bw.cuh with TILE-LOCAL STAGING of the gather, the companion of
// spiral_conv_umma_staged.cuh (same plan: plan_cnt [L], plan_src [L, rcap/2], plan_loc [L, S, 128]).  Written at
// the end of round 1 after the GPU budget was spent: compiled for sm_100a, never run.  DESIGN.md section 7 item 1.
//
// Differences from bw_umma_kernel: one loader warp per tile stage copies the tile's DISTINCT source rows once
// (instead of 128 gathered rows per slot); the transposing splitter of (slot s, channel c, row group r4) reads
// its channel of staged rows loc[tile, s, 32*r4 + j] -- each read is still one whole 128-byte staged row per
// warp instruction, so arbitrary rows cost no bank conflicts; the positions reach the lanes by shuffle from one
// coalesced load.  g staging, the MMA schedule, the accumulator drains and the partial layout are unchanged, so
// the result is bit-identical to bw_umma_kernel's.
// Barriers: tile_full[ts] (32 arrivals of the loader warp), tile_empty[ts] (32 threads x 4 row groups x S slots).
#pragma once
#include "spiral_conv_umma_bw.cuh"
#include "spiral_conv_umma_staged.cuh"

namespace sdvae {
namespace umma {

struct BwStagedArgs {
    const float* in;          // [B, in_rows, in_ld], the 32 channels of this pass start at `in`
    const int* plan_cnt;      // [L]
    const int* plan_src;      // [L, rcap/2]
    const int* plan_loc;      // [L, S, 128]
    const float* g;           // [B, out_rows, g_ld]
    float* part;              // per-CTA partial dW (layout of BwUmmaArgs::part)
    float* part_b;            // per-CTA partial db or nullptr
    int in_ld, g_ld, part_ld, part_cta, partb_cta;
    int B, in_rows, out_rows, L, S, rcap, n_real, nts;
    int flush;                // tiles per accumulator drain (>= 1)
};

__global__ void __launch_bounds__(kBwThreads, 1)
bw_umma_staged_kernel(const BwStagedArgs a) {
    const int S = a.S;
    const int NTS = a.nts;                                // tile-stage ring depth
    const int TILE_STAGE = a.rcap * 128;
    const int NBLK = (S * 32 + 1 + 127) >> 7;             // accumulator blocks (incl. the ones row)
    const int ONES_ROW = S * 32;                          // M row of the db accumulator
    const int K = S * 32;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* G_s = smem;                                   // [2][16][hi 1 KB | lo 1 KB]
    uint8_t* R_s = G_s + 2 * kGStage;                      // [NTS][rcap][128 B]  distinct source rows of a tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(R_s + (size_t)NTS * TILE_STAGE);
    uint64_t* tile_full = bars;                            // [NTS] loader warp (32 arrivals) -> splitters
    uint64_t* tile_empty = bars + kTileStages;             // [NTS] splitters (S slots x 4 row groups x 32 threads) -> loader
    uint64_t* a_full = bars + 2 * kTileStages;
    uint64_t* a_empty = a_full + kBwAStages;
    uint64_t* g_full = a_empty + kBwAStages;                 // [2] g warps -> MMA
    uint64_t* g_empty = g_full + 2;                        // [2] MMA (commit) -> g warps
    uint64_t* done_bar = g_empty + 2;                      // [2] MMA (commit) -> drain, per accumulator set
    uint64_t* drained_bar = done_bar + 2;                  // [2] drain -> MMA: the set may be restarted
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained_bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kTileStages; ++i) { mbar_init(tile_full + i, 32); mbar_init(tile_empty + i, 128 * S); }
        for (int i = 0; i < kBwAStages; ++i) { mbar_init(a_full + i, 128); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(g_full + i, 128); mbar_init(g_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(done_bar + i, 1); mbar_init(drained_bar + i, kBwEpilogueWarps * 32); }
        fence_barrier_init();
    }
    if (warp == kBwMmaWarp) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = a.B * a.L;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int CPT = NBLK * 4;                             // A chunks per tile: (block, 32-row group)
    const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;
    const int T = a.flush;

    if (warp == kBwMmaWarp) {
        // ================= MMA issuer =================
        constexpr uint32_t IDESC = idesc_tf32_bmn(kBM, kBwNT);
        const bool leader = elect_one();
        const uint32_t g_base = smem_u32(G_s);
        int as = 0; uint32_t aph = 0;
#pragma unroll 1
        int tf = 0, nfl = 0;                               // tile index inside the flush group, flushes done
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int gb = it & 1;
            const int ab = nfl & 1;                        // accumulator set of this flush group
            if (tf == 0 && nfl >= 2) {                     // previous use of this set drained?
                mbar_wait(drained_bar + ab, (uint32_t)(((nfl >> 1) - 1) & 1));
                tc_fence_after();
            }
            const bool last_of_group = (tf == T - 1) || (it == my_tiles - 1);
#pragma unroll 1
            for (int c = 0; c < CPT; ++c) {
                const int blk = c >> 2, r4 = c & 3;
                mbar_wait(a_full + as, aph);
                if (c == 0) mbar_wait(g_full + gb, (it >> 1) & 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)((ab * NBLK + blk) * kBwNT);
                    const uint32_t a_hi = tmem_base + (uint32_t)(kBwAColBase + as * 64), a_lo = a_hi + 32;
                    const uint32_t g_t = g_base + gb * kGStage + r4 * 4 * 2048;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd_hi = smem_desc_mn_sw128(g_t + k * 2048);
                        const uint64_t bd_lo = smem_desc_mn_sw128(g_t + k * 2048 + 1024);
                        umma_tf32_ts(d_tmem, a_hi + k * 8, bd_hi, IDESC, (tf | r4 | k) != 0);
                        umma_tf32_ts(d_tmem, a_hi + k * 8, bd_lo, IDESC, 1u);
                        umma_tf32_ts(d_tmem, a_lo + k * 8, bd_hi, IDESC, 1u);
                    }
                    umma_commit(a_empty + as);
                    if (c == CPT - 1) {
                        umma_commit(g_empty + gb);
                        if (last_of_group) umma_commit(done_bar + ab);
                    }
                }
                __syncwarp();
                if (++as == kBwAStages) { as = 0; aph ^= 1; }
            }
            if (last_of_group) { tf = 0; ++nfl; } else ++tf;
        }
    } else if (warp < kBwFirstSplitWarp) {
        // ================= g staging (per tile), then the epilogue =================
        const int p = tid;                                        // 0..127
        int b = (int)blockIdx.x / a.L, jt = (int)blockIdx.x - b * a.L;
        const int n_real = a.n_real;
        const bool vec = (n_real == kBwNT) && ((reinterpret_cast<uintptr_t>(a.g) & 15) == 0) && (a.g_ld % 4 == 0);
        // ---- drain f: D[(s,c), n] += into the partial dW[n, s*32 + c], ones row -> partial db[n] ----
        const int q4 = warp & 3;
        float* P = a.part + (size_t)blockIdx.x * a.part_cta;
        float* Pb = a.part_b ? a.part_b + (size_t)blockIdx.x * a.partb_cta : nullptr;
        auto drain = [&](int f) {
            const int ab = f & 1;
            mbar_wait(done_bar + ab, (uint32_t)((f >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int blk = 0; blk < NBLK; ++blk) {
                const int row = blk * 128 + q4 * 32 + lane;       // M row = s*32 + c
                const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((ab * NBLK + blk) * kBwNT);
                float v1[16], v2[16];
                tmem_ld16(t_row, v1);
                tmem_ld16(t_row + 16, v2);
                tmem_ld_wait();
                if (blk == NBLK - 1) {                            // last read of this set: hand it back
                    tc_fence_before();
                    mbar_arrive(drained_bar + ab);
                }
                // M row = s*32 + c -> column s*in_ld + c of the layer's [n, S*C_in] weight gradient
                float* dst = row < K ? P + (row >> 5) * a.in_ld + (row & 31) : (row == ONES_ROW ? Pb : nullptr);
                const size_t ld = row < K ? (size_t)a.part_ld : (size_t)1;
                if (dst) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j < n_real) atomicAdd(dst + (size_t)j * ld, v1[j]);          // RED, single writer
                        if (16 + j < n_real) atomicAdd(dst + (size_t)(16 + j) * ld, v2[j]);
                    }
                }
            }
        };
        int tf = 0, nfl = 0;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int gb = it & 1;
            uint8_t* gs = G_s + gb * kGStage;
            const int nvalid = min(kBM, a.out_rows - jt * kBM);
            const float* gt = a.g + ((size_t)b * a.out_rows + (size_t)jt * kBM) * a.g_ld;
            float4 v[8];
            if (vec) {                                            // 8 lanes per row, coalesced 16-byte loads
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int m = (p >> 3) + 16 * i;
                    v[i] = m < nvalid ? ldg4(gt + (size_t)m * a.g_ld + 4 * (p & 7)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {                                              // narrow rows: thread = row, scalar loads
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float t[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = 4 * i + j;
                        t[j] = (p < nvalid && n < n_real) ? __ldg(gt + (size_t)p * a.g_ld + n) : 0.f;
                    }
                    v[i] = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            mbar_wait(g_empty + gb, ((it >> 1) & 1) ^ 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = vec ? (p >> 3) + 16 * i : p;
                const int q = vec ? (p & 7) : i;
                float4 hi, lo;
                split_tf32f(v[i].x, hi.x, lo.x); split_tf32f(v[i].y, hi.y, lo.y);
                split_tf32f(v[i].z, hi.z, lo.z); split_tf32f(v[i].w, hi.w, lo.w);
                uint8_t* dst = gs + g_off(m, q);
                *reinterpret_cast<float4*>(dst) = hi;
                *reinterpret_cast<float4*>(dst + 1024) = lo;
            }
            fence_async_smem();
            mbar_arrive(g_full + gb);
            b += db; jt += djt;
            if (jt >= a.L) { jt -= a.L; ++b; }
            // tile `it` is staged; if tile it-1 closed a flush group, drain it now (the MMA warp is waiting)
            if (tf == 0 && it > 0) { drain(nfl); ++nfl; }
            tf = (tf == T - 1) ? 0 : tf + 1;
        }
        if (my_tiles > 0) drain(nfl);                             // the last group always ends with a flush
    } else if (warp < kBwFirstLoadWarp) {
        // ================= splitters =================
        const int set = (warp - kBwFirstSplitWarp) >> 2;            // set 0: even 32-row groups, set 1: odd
        const int q4 = warp & 3;
        const int sw_lane = lane >> 2, w_lane = (lane & 3) * 4;
        int jt = (int)blockIdx.x % a.L;
        int as = set; uint32_t aph = 0;                            // A stage of chunk c = it*CPT + blk*4 + r4
        int ts = 0; uint32_t tph = 0;                              // tile stage / phase of tile iteration it
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int nvalid = min(kBM, a.out_rows - jt * kBM);
#pragma unroll 1
            for (int c = set; c < CPT; c += 2) {
                const int blk = c >> 2, r4 = c & 3;
                const int s = blk * 4 + q4;
                const int row0 = blk * 128 + q4 * 32;              // first M row of this warp
                mbar_wait(a_empty + as, aph ^ 1);
                float v[32];
                if (s < S) {
                    // lane j holds the staged-row position of tile row 32*r4 + j for this slot (0 past the mesh)
                    const int locv = __ldg(a.plan_loc + ((size_t)jt * S + s) * kBM + 32 * r4 + lane);
                    mbar_wait(tile_full + ts, tph);
                    const uint8_t* stage = R_s + (size_t)ts * TILE_STAGE + w_lane;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int e = __shfl_sync(0xffffffffu, locv, j);          // warp-uniform staged row
                        v[j] = 32 * r4 + j < nvalid ? *reinterpret_cast<const float*>(stage + e * 128 + ((sw_lane ^ (e & 7)) << 4)) : 0.f;
                    }
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { float h; split_tf32f(v[j], h, lo[j]); v[j] = h; }
                    tc_fence_after();
                    const uint32_t t_a = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(kBwAColBase + as * 64);
                    tmem_st32(t_a, v);
                    tmem_st32(t_a + 32, lo);
                    mbar_arrive(tile_empty + ts);
                    tmem_st_wait();
                } else if (row0 <= ONES_ROW && ONES_ROW < row0 + 32) {
                    // the warp that owns the ones row: A^T[ONES_ROW, m] = 1 for valid rows, everything else 0
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] = (row0 + lane == ONES_ROW && 32 * r4 + j < nvalid) ? 1.f : 0.f;
                        lo[j] = 0.f;
                    }
                    tc_fence_after();
                    const uint32_t t_a = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(kBwAColBase + as * 64);
                    tmem_st32(t_a, v);
                    tmem_st32(t_a + 32, lo);
                    tmem_st_wait();
                }
                // (warps past the ones row leave their TMEM lanes alone: those accumulator rows are never read)
                tc_fence_before();
                mbar_arrive(a_full + as);
                as += 2;
                if (as >= kBwAStages) { as -= kBwAStages; aph ^= 1; }
            }
            // CPT is a multiple of 4 = kBwAStages, so the (as, aph) sequence continues seamlessly into the next tile
            if (++ts == NTS) { ts = 0; tph ^= 1; }
            jt += djt; if (jt >= a.L) jt -= a.L;
        }
    } else {
        // ================= loaders: warp lw owns tile stage lw, one whole tile per pass =================
        const int lw = warp - kBwFirstLoadWarp;
        if (lw < NTS) {
            const int q = lane & 7, rsub = lane >> 3;
            const uint32_t sw0 = (uint32_t)((q ^ rsub) << 4), sw1 = (uint32_t)((q ^ (rsub + 4)) << 4);
            const uint32_t dst = smem_u32(R_s) + (uint32_t)lw * (uint32_t)TILE_STAGE + (uint32_t)rsub * 128u;
            constexpr int PV = kStagedMaxRcap / 32;
            long long t0 = (long long)blockIdx.x + (long long)lw * gridDim.x;       // global tile of iteration lw
            int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
            uint32_t tph = 0;
#pragma unroll 1
            for (int it = lw; it < my_tiles; it += NTS) {
                PlanRegs<PV> now;
                plan_fetch(now, a.plan_cnt, a.plan_src, jt, 1, 0, a.rcap, rsub);
                const float* base = a.in + (size_t)b * a.in_rows * a.in_ld + 4 * q;
                mbar_wait(tile_empty + lw, tph ^ 1);
                plan_issue(now, dst + sw0, dst + sw1, base, (uint32_t)a.in_ld * 4u);
                cp_async_commit();
                cp_async_wait<0>();
                mbar_arrive(tile_full + lw);
                tph ^= 1;
                for (int k = 0; k < NTS; ++k) {                 // advance NTS tiles of this CTA's schedule
                    b += db; jt += djt;
                    if (jt >= a.L) { jt -= a.L; ++b; }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kBwMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace umma
}  // namespace sdvae
