// SpiralConv weight gradient on the tensor cores (tcgen05 + TMEM), C_in = 32 per slot.
// Reference: autograd of nn.Linear in model.py:40 over the (never materialised) gather of model.py:34.
//
//   dW[n, s*32 + c] = sum_m g[m, n] * x[src(m, s), c]          db[n] = sum_m g[m, n]
//
// As an MMA:  D[(s,c), n] += A^T[(s,c), m] * g[m, n]  with the mesh rows m as the contraction dimension.
//   * M = 128 = four spiral slots x 32 channels per accumulator block; ceil((32*S + 1) / 128) blocks; the
//     first unused M row is a row of ones, which makes the tensor core produce db as well.
//   * A^T comes from TMEM (TS form): the splitter thread of (slot, channel c) reads its channel of 32
//     consecutive staged rows (LDS.32, one 128-byte row per warp instruction), splits hi/lo and writes
//     32 + 32 TMEM columns -- the transpose is free, it is just which thread reads what.
//   * g is the shared-memory operand, MN-major (n contiguous), in the 32-byte-base 128B swizzle that
//     MN-major tf32 requires; the hi image and the lo image of the same 8 rows sit 1024 B apart.
//     3xTF32 = three N = 32 MMAs per k-step onto ONE 32-column accumulator:
//     A_hi^T g_hi + A_hi^T g_lo + A_lo^T g_hi  (one accumulator per block leaves TMEM room for two sets).
//   * the tensor core adds into its fp32 accumulator with truncation, so a long accumulation chain drifts
//     (measured: 2.8e-5 normwise after 58 tiles).  The accumulators are therefore DRAINED every `flush`
//     tiles: the epilogue warps read them out of TMEM and add them (round-to-nearest, RED.ADD.F32, one
//     writer per address in program order => deterministic) into the CTA's own partial dW / db in L2.
//     Two accumulator sets alternate between flush groups, so the drain of group f overlaps the MMAs of
//     group f + 1.  split_reduce_kernel adds the per-CTA partials in CTA order.
//
// Warp roles (640 threads): warps 0..3 stage g (load, split, store, proxy fence) and run the final
// epilogue; warps 4..11 splitters (two sets alternating A chunks); warps 12..18 cp.async loaders
// (same raw ring and tile plan as the forward kernel); warp 19 TMEM allocation + MMA issue.
#pragma once
#include "spiral_conv_umma.cuh"

namespace sdvae {
namespace umma {

constexpr int kBwNT = 32;                 // output channels per tile (n_real <= 32)
// warp roles of this kernel (20 warps): 0..3 g staging + drain, 4..11 splitters, 12..18 loaders, 19 MMA
constexpr int kBwEpilogueWarps = 4;
constexpr int kBwFirstSplitWarp = 4;
constexpr int kBwFirstLoadWarp = 12;
constexpr int kBwMmaWarp = 19;
constexpr int kBwThreads = (kBwMmaWarp + 1) * 32;
constexpr int kBwAStages = 4;             // TMEM A-operand ring (64 columns each: 32 hi + 32 lo)
constexpr int kBwAColBase = 256;          // TMEM columns [0,192): two accumulator sets, [256,512): A ring
constexpr int kGStage = 16 * 2048;        // 16 K-atoms x (hi atom + lo atom)

struct BwUmmaArgs {
    const float* in;          // [B, in_rows, in_ld], the 32 channels of this pass start at `in`
    const int* plan_cnt;      // [L, S]        forward tile plan: one staged row per tile row, in row order
    const int* plan_src;      // [L, S, rcap/2]  packed (plan_fetch)
    const float* g;           // [B, out_rows, g_ld]     gradient w.r.t. the pre-activation, n_real columns from `g`
    float* part;              // per-CTA partial dW, zero-initialised by the caller: element (n, s, c) of CTA k at
                              //   part[k*part_cta + n*part_ld + s*in_ld + c]   (the caller offsets `part` to the
                              //   first output channel / input channel of this pass)
    float* part_b;            // per-CTA partial db at part_b[k*partb_cta + n], or nullptr (no bias from this pass)
    int in_ld, g_ld, part_ld, part_cta, partb_cta;
    int B, in_rows, out_rows, L, S, rcap, n_real, nraw;
    int flush;                // tiles per accumulator drain (>= 1)
};

// MN-major tf32 operands have one legal shared-memory layout: SWIZZLE_128B with a 32-byte base
// (layout type 1, Swizzle<2,5,2>): atoms of 4 K-rows x 128 B (32 floats along N); inside an atom the
// 32-byte column j of row r sits at r*128 + ((j ^ r) * 32).  One K = 8 MMA reads two K-atoms (SBO apart);
// the next 32 floats along N (here: the lo image of the same rows) are LBO apart.
//   per 8 mesh rows: [hi rows 0-3 | hi rows 4-7 | lo rows 0-3 | lo rows 4-7], 512 B each
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of the 16-byte piece q (0..7) of mesh row m (0..127) in the hi image of a g stage
__device__ __forceinline__ int g_off(int m, int q) {
    return (m >> 3) * 2048 + ((m >> 2) & 1) * 512 + (m & 3) * 128 + ((((q >> 1) ^ (m & 3))) << 5) + (q & 1) * 16;
}
// D = F32, A = B = TF32, A K-major (TMEM), B MN-major
__host__ __device__ constexpr uint32_t idesc_tf32_bmn(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kBwThreads, 1)
bw_umma_kernel(const BwUmmaArgs a) {
    const int S = a.S;
    const int NRAW = a.nraw;
    const int RAW_STAGE = a.rcap * 128;
    const int NBLK = (S * 32 + 1 + 127) >> 7;             // accumulator blocks (incl. the ones row)
    const int ONES_ROW = S * 32;                          // M row of the db accumulator
    const int K = S * 32;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* G_s = smem;                                   // [2][16][hi 1 KB | lo 1 KB]
    uint8_t* R_s = G_s + 2 * kGStage;                      // [NRAW][rcap][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(R_s + (size_t)NRAW * RAW_STAGE);
    uint64_t* raw_full = bars;
    uint64_t* raw_empty = bars + kMaxRaw;
    uint64_t* a_full = bars + 2 * kMaxRaw;
    uint64_t* a_empty = a_full + kBwAStages;
    uint64_t* g_full = a_empty + kBwAStages;                 // [2] g warps -> MMA
    uint64_t* g_empty = g_full + 2;                        // [2] MMA (commit) -> g warps
    uint64_t* done_bar = g_empty + 2;                      // [2] MMA (commit) -> drain, per accumulator set
    uint64_t* drained_bar = done_bar + 2;                  // [2] drain -> MMA: the set may be restarted
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained_bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NRAW; ++i) { mbar_init(raw_full + i, 32); mbar_init(raw_empty + i, 128); }
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
        int rs0 = 0; uint32_t rph0 = 0;                            // raw stage / phase of (tile it, slot 0)
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
                    int rs = rs0 + s; uint32_t rph = rph0;
                    while (rs >= NRAW) { rs -= NRAW; rph ^= 1; }
                    mbar_wait(raw_full + rs, rph);
                    const uint8_t* stage = R_s + (size_t)rs * RAW_STAGE + w_lane;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int e = 32 * r4 + j;                 // staged row = tile row (forward plan)
                        // rows past nvalid were never staged (and lie outside the stage when rcap < 128)
                        v[j] = e < nvalid ? *reinterpret_cast<const float*>(stage + e * 128 + ((sw_lane ^ (e & 7)) << 4)) : 0.f;
                    }
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { float h; split_tf32f(v[j], h, lo[j]); v[j] = h; }
                    tc_fence_after();
                    const uint32_t t_a = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(kBwAColBase + as * 64);
                    tmem_st32(t_a, v);
                    tmem_st32(t_a + 32, lo);
                    mbar_arrive(raw_empty + rs);
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
            rs0 += S;
            while (rs0 >= NRAW) { rs0 -= NRAW; rph0 ^= 1; }
            jt += djt; if (jt >= a.L) jt -= a.L;
        }
    } else {
        // ================= loaders (raw stage w owned by loader warp w; chunk = (tile, slot)) =================
        const int lw = warp - kBwFirstLoadWarp;
        const int q = lane & 7, rsub = lane >> 3;
        const uint32_t sw0 = (uint32_t)((q ^ rsub) << 4), sw1 = (uint32_t)((q ^ (rsub + 4)) << 4);
        const uint32_t raw_base = smem_u32(R_s) + (uint32_t)rsub * 128u;
        const int G = my_tiles * S;
        // position of chunk g = lw: tile iteration g / S, slot g % S
        int g = lw, sl = lw, b = (int)blockIdx.x / a.L, jt = (int)blockIdx.x - b * a.L;
        while (sl >= S) { sl -= S; b += db; jt += djt; if (jt >= a.L) { jt -= a.L; ++b; } }
        uint32_t rph = 0;
        if (lw >= NRAW) g = G;
        PlanRegs<4> nxt;                                           // forward plan: <= 128 staged rows
        if (g < G) plan_fetch(nxt, a.plan_cnt, a.plan_src, jt, S, sl, a.rcap, rsub);
#pragma unroll 1
        while (g < G) {
            const PlanRegs<4> now = nxt;
            const float* base = a.in + (size_t)b * a.in_rows * a.in_ld + 4 * q;
            g += NRAW; sl += NRAW;
            while (sl >= S) { sl -= S; b += db; jt += djt; if (jt >= a.L) { jt -= a.L; ++b; } }
            if (g < G) plan_fetch(nxt, a.plan_cnt, a.plan_src, jt, S, sl, a.rcap, rsub);
            mbar_wait(raw_empty + lw, rph ^ 1);
            const uint32_t dst = raw_base + (uint32_t)lw * (uint32_t)RAW_STAGE;
            plan_issue(now, dst + sw0, dst + sw1, base, (uint32_t)a.in_ld * 4u);
            cp_async_commit();
            cp_async_wait<0>();
            mbar_arrive(raw_full + lw);
            rph ^= 1;
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
