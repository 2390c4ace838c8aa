// EXPERIMENTAL (compiled, NOT dispatched by the engine, the autograd functions or bench.py; reached only through
// sdvae_spiralconv_fwd_tc_staged, which tests/ exercise only when SDVAE_EXPERIMENTAL=1): the tcgen05 SpiralConv
// forward of spiral_conv_umma.cuh with TILE-LOCAL STAGING of the gather.  Written at the end of round 1 after the
// GPU budget was spent -- it has been compiled for sm_100a but never run; DESIGN.md section 7 item 1 is its plan.
//
// gc_umma_kernel copies, for every tile of 128 output rows and every spiral slot, the 128 gathered rows
// (9 x 128 x 128 B per tile): it is bound by the L2->SM fabric (~7-8 TB/s of gathered rows), not by HBM.
// A tile reads far fewer DISTINCT rows (412 in the template's strip order, ~200 after tables.patch_order), so
// here ONE loader warp copies the tile's distinct rows once into a tile stage (ring of kTileStages), and the
// splitter of tile row r reads staged row loc[tile, s, r] for slot s.  Everything downstream of the splitter
// (TMEM A ring, the MMA-issuing thread, the epilogue, the weight image) is gc_umma_kernel's.
//
//   plan_cnt [L]            distinct source rows of tile t
//   plan_src [L, rcap/2]    those rows, 16-bit pairs in loader-lane order (umma::plan_fetch with S = 1)
//   plan_loc [L, S, 128]    position of idx[t*128 + r, s] in tile t's list (0 for rows past the mesh)
//
// Barriers: tile_full[ts] (the loader warp, count 1), tile_empty[ts] (every splitter warp after every chunk it
// read: 4 * S arrivals per tile).  A splitter waits for a_empty (its TMEM stage) first, then for tile_full, as in
// gc_umma_kernel; the loader of tile i + kTileStages cannot refill a stage before all 4 * S reads of tile i have
// arrived, so no waiter is ever more than one phase behind its barrier.
#pragma once
#include "spiral_conv_umma.cuh"

namespace sdvae {
namespace umma {

constexpr int kTileStages = 3;            // tile-stage ring depth limit (one loader warp per stage)
constexpr int kStagedMaxRcap = 288;       // distinct rows per tile the kernel supports (multiple of 32)

struct StagedArgs {
    const float* in;          // [B, in_rows, 32]
    const int* plan_cnt;      // [L]
    const int* plan_src;      // [L, rcap/2]
    const int* plan_loc;      // [L, S, 128]
    const float* wimg;        // packed weight image (umma_pack_weights_kernel, KS = 32)
    const float* bias;        // [n_real] or nullptr
    float* out;               // [B, out_rows, ldo]
    int B, in_rows, out_rows, L, S, rcap;
    int n_real, ldo, epi;     // epi: EPI_BIAS or EPI_BIAS_ELU
    int nts;                  // tile-stage ring depth (2 .. kTileStages)
    int ostage;               // 1: epilogue stages output rows in shared memory (NT == 32 only)
};

template <int NT>
struct StagedCfg {
    static constexpr int B_CHUNK = 2 * NT * 128;
    static constexpr int ACC_COLS = 4 * NT;
    static constexpr int MAX_AST = (kTmemCols - ACC_COLS) / 64 < kMaxAStages ? (kTmemCols - ACC_COLS) / 64 : kMaxAStages;
    static size_t b_bytes(int S) { return (size_t)S * B_CHUNK; }
    static bool out_stage() { return NT == 32; }
    static int tile_stages(int S, int rcap) {
        const long long budget = 226LL * 1024 - 2048 - (long long)b_bytes(S) - (out_stage() ? kOutStageBytes : 0);
        long long st = budget / ((long long)rcap * 128);
        return (int)(st > kTileStages ? kTileStages : st);
    }
    static size_t smem_bytes(int S, int rcap, int nts) {
        return 1024 + b_bytes(S) + (size_t)nts * rcap * 128 + (out_stage() ? kOutStageBytes : 0) + 1024;
    }
};

// position of a splitter set in the CTA's schedule: chunk g = (tile iteration ti, slot ch)
struct StagedCursor {
    int g, ch, b, jt, as, ts;
    uint32_t aph, tph;
    int nch, nast, nts, L, db, djt;
    __device__ __forceinline__ StagedCursor(const StagedArgs& a, int nast_)
        : g(0), ch(0), as(0), ts(0), aph(0), tph(0), nch(a.S), nast(nast_), nts(a.nts), L(a.L) {
        b = (int)blockIdx.x / L; jt = (int)blockIdx.x - b * L;
        db = (int)gridDim.x / L; djt = (int)gridDim.x - db * L;
    }
    __device__ __forceinline__ void advance(int step) {
        g += step;
        as += step; while (as >= nast) { as -= nast; aph ^= 1; }
        ch += step;
        while (ch >= nch) {
            ch -= nch; b += db; jt += djt;
            if (jt >= L) { jt -= L; ++b; }
            if (++ts == nts) { ts = 0; tph ^= 1; }
        }
    }
};

template <int NT>
__global__ void __launch_bounds__(kThreads, 1)
gc_umma_staged_kernel(const StagedArgs a) {
    using Cfg = StagedCfg<NT>;
    constexpr int B_CHUNK = Cfg::B_CHUNK, ACOL = Cfg::ACC_COLS;
    const int S = a.S;
    const int NCH = S;                                      // KS = 32: one 32-wide chunk per slot
    const int NTS = a.nts;
    const int NAST = Cfg::MAX_AST;
    const int NS = NAST < kSplitSets ? NAST : kSplitSets;
    const int TILE_STAGE = a.rcap * 128;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* B_s = smem;                                    // [S][2NT][128 B]     resident weight image
    uint8_t* R_s = B_s + (size_t)NCH * B_CHUNK;             // [NTS][rcap][128 B]  distinct source rows of a tile
    uint8_t* O_s = R_s + (size_t)NTS * TILE_STAGE;          // [128][144 B]        output staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(O_s + (a.ostage ? kOutStageBytes : 0));
    uint64_t* tile_full = bars;                             // [NTS]  loader    -> splitters
    uint64_t* tile_empty = bars + kTileStages;              // [NTS]  splitters -> loader
    uint64_t* a_full = bars + 2 * kTileStages;              // [NAST] splitters -> MMA
    uint64_t* a_empty = a_full + kMaxAStages;               // [NAST] MMA (commit) -> splitters
    uint64_t* t_full = a_empty + kMaxAStages;               // [2]    MMA (commit) -> epilogue
    uint64_t* t_empty = t_full + 2;                         // [2]    epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kTileStages; ++i) { mbar_init(tile_full + i, 1); mbar_init(tile_empty + i, 4 * NCH); }
        for (int i = 0; i < kMaxAStages; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, kEpilogueWarps); }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    {
        const int n16 = NCH * B_CHUNK / 16;
        const float4* src = reinterpret_cast<const float4*>(a.wimg);
        float4* dst = reinterpret_cast<float4*>(B_s);
#pragma unroll 1
        for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = a.B * a.L;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int G = my_tiles * NCH;

    if (warp >= kFirstLoadWarp) {
        reg_dec<kRegsLoad>();
        if (warp == kMmaWarp) {
            // ================= MMA issuer (as gc_umma_kernel) =================
            if (elect_one()) {
                constexpr uint32_t IDESC1 = idesc_tf32(kBM, 2 * NT);
                constexpr uint32_t IDESC2 = idesc_tf32(kBM, NT);
                const uint64_t desc0 = smem_desc_sw128(smem_u32(B_s));
                const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
                int as = 0; uint32_t aph = 0;
                bool ready = my_tiles > 0 && mbar_try_wait(a_full, 0u);
#pragma unroll 1
                for (int it = 0; it < my_tiles; ++it) {
                    const int acc = it & 1;
                    mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * NT);
#pragma unroll 1
                    for (int ch = 0; ch < NCH; ++ch) {
                        if (!ready) mbar_wait(a_full + as, aph);
                        tc_fence_after();
                        const uint32_t a_hi = tmem_base + (uint32_t)(ACOL + as * 64), a_lo = a_hi + 32;
                        const uint32_t dl = desc_lo0 + (uint32_t)(ch * (B_CHUNK >> 4));
                        uint64_t* const my_empty = a_empty + as;
                        if (++as == NAST) { as = 0; aph ^= 1; }
                        ready = (ch + 1 < NCH || it + 1 < my_tiles) && mbar_try_wait(a_full + as, aph);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(dl + 2u * k);
                            umma_tf32_ts(d_tmem, a_hi + k * 8, bd, IDESC1, (ch | k) != 0);
                            umma_tf32_ts(d_tmem, a_lo + k * 8, bd, IDESC2, 1u);
                        }
                        umma_commit(my_empty);
                        if (ch == NCH - 1) umma_commit(t_full + acc);
                    }
                }
            }
            __syncwarp();
        } else {
            // ================= loaders: warp lw owns tile stage lw, one whole tile per pass =================
            const int lw = warp - kFirstLoadWarp;
            if (lw < NTS) {
                const int q = lane & 7, rsub = lane >> 3;
                const uint32_t sw0 = (uint32_t)((q ^ rsub) << 4), sw1 = (uint32_t)((q ^ (rsub + 4)) << 4);
                const uint32_t dst = smem_u32(R_s) + (uint32_t)lw * (uint32_t)TILE_STAGE + (uint32_t)rsub * 128u;
                constexpr int PV = kStagedMaxRcap / 32;
                const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;
                // tile iteration `it` of this CTA is global tile blockIdx.x + it * gridDim.x
                long long t0 = (long long)blockIdx.x + (long long)lw * gridDim.x;
                int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
                uint32_t tph = 0;
#pragma unroll 1
                for (int it = lw; it < my_tiles; it += NTS) {
                    PlanRegs<PV> now;
                    plan_fetch(now, a.plan_cnt, a.plan_src, jt, 1, 0, a.rcap, rsub);
                    const float* base = a.in + (size_t)b * a.in_rows * 32 + 4 * q;
                    mbar_wait_relaxed(tile_empty + lw, tph ^ 1);
                    plan_issue(now, dst + sw0, dst + sw1, base, 128u);
                    cp_async_commit();
                    cp_async_wait<0>();
                    warp_arrive(tile_full + lw, lane);
                    tph ^= 1;
                    for (int k = 0; k < NTS; ++k) {                 // advance NTS tiles of this CTA's schedule
                        b += db; jt += djt;
                        if (jt >= a.L) { jt -= a.L; ++b; }
                    }
                }
            }
        }
    } else if (warp < kFirstSplitWarp) {
        reg_dec<kRegsEpilogue>();
        // ================= epilogue (as gc_umma_kernel, forward epilogues only) =================
        const int q4 = warp & 3;
        const int EPI = a.epi;
        const int n_real = a.n_real, ldo = a.ldo;
        const bool has_bias = a.bias != nullptr;
        constexpr bool vec_ok = NT >= 32;
        constexpr bool kCanStage = NT == 32;
        int b = (int)blockIdx.x / a.L, jt = (int)blockIdx.x - b * a.L;
        const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            mbar_wait_relaxed(t_full + acc, (it >> 1) & 1);
            tc_fence_after();
            const int r = jt * kBM + q4 * 32 + lane;
            const bool row_ok = r < a.out_rows;
            const size_t m = (size_t)b * a.out_rows + r;
            const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * 2 * NT);
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 16) {
                float v[16], d2[16];
                tmem_ld16(t_row + c0, v);
                tmem_ld16(t_row + NT + c0, d2);
                tmem_ld_wait();
                if (c0 + 16 >= NT) {
                    tc_fence_before();
                    warp_arrive(t_empty + acc, lane);
                }
                if (!row_ok) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += d2[j];
                float* orow = a.out + m * ldo + c0;
                if (has_bias) {
                    if (vec_ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 bv = ldg4(a.bias + c0 + j);
                            v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < n_real) v[j] += __ldg(a.bias + c0 + j);
                    }
                }
                if (EPI == EPI_BIAS_ELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = elu_fast(v[j]);
                }
                if (vec_ok) {
                    if (kCanStage && a.ostage) {
                        float4* srow = reinterpret_cast<float4*>(O_s + (q4 * 32 + lane) * kOutRowBytes + c0 * 4);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) srow[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < n_real) orow[j] = v[j];
                }
            }
            if (kCanStage && a.ostage) {
                __syncwarp();
                const int piece = lane & 7;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int lr2 = q4 * 32 + 4 * k + (lane >> 3);
                    const float4 t = *reinterpret_cast<const float4*>(O_s + lr2 * kOutRowBytes + piece * 16);
                    const int r2 = jt * kBM + lr2;
                    if (r2 < a.out_rows)
                        *reinterpret_cast<float4*>(a.out + ((size_t)b * a.out_rows + r2) * ldo + piece * 4) = t;
                }
                __syncwarp();
            }
            b += db; jt += djt;
            if (jt >= a.L) { jt -= a.L; ++b; }
        }
    } else {
        reg_inc<kRegsSplit>();
        // ================= splitters: tile row lr reads staged row loc[tile, slot, lr] =================
        const int set = (warp - kFirstSplitWarp) >> 2;
        const int q4 = warp & 3;
        const int lr = q4 * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)ACOL;
        if (set < NS) {
            StagedCursor cur(a, NAST);
            cur.advance(set);
            int loc = 0;
            if (cur.g < G) loc = __ldg(a.plan_loc + ((size_t)cur.jt * S + cur.ch) * kBM + lr);
#pragma unroll 1
            while (cur.g < G) {
                const int as = cur.as, ts = cur.ts;
                const uint32_t aph = cur.aph, tph = cur.tph;
                cur.advance(NS);
                int loc_next = 0;                                 // one own-chunk ahead
                if (cur.g < G) loc_next = __ldg(a.plan_loc + ((size_t)cur.jt * S + cur.ch) * kBM + lr);
                mbar_wait_relaxed(a_empty + as, aph ^ 1);        // order: a_empty first, then the tile stage
                mbar_wait_relaxed(tile_full + ts, tph);
                const uint8_t* row = R_s + (size_t)ts * TILE_STAGE + (size_t)loc * 128;
                const int x7 = loc & 7;
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 t = *reinterpret_cast<const float4*>(row + ((j ^ x7) << 4));
                    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
                }
                float lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) { float h; split_tf32f(v[j], h, lo[j]); v[j] = h; }
                tc_fence_after();
                const uint32_t t_a = t_lane + (uint32_t)(as * 64);
                tmem_st32(t_a, v);
                tmem_st32(t_a + 32, lo);
                warp_arrive(tile_empty + ts, lane);    // this warp's reads of the tile stage for this chunk are done
                tmem_st_wait();
                tc_fence_before();
                warp_arrive(a_full + as, lane);
                loc = loc_next;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace umma
}  // namespace sdvae
