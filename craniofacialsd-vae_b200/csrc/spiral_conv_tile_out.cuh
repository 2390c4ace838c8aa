// Narrow-OUTPUT spiral convolution forward (32 -> NO channels, S*NO <= 32: the 32 -> 3 output layer, model.py:135-136,
// 172 through model.py:27-41) on tcgen05 by PROJECT-THEN-GATHER:
//
//   y[v, n] = bias[n] + sum_s P[src(v, s), s*NO + n],        P[u, s*NO + n] = sum_c W[n, s*32 + c] * x[u, c]
//
// The gathered A operand of the generic kernel (spiral_conv_tile.cuh) costs 9 x 32 KB of TMEM stores per 128-row tile
// -- for 3 useful output channels.  Here the tile's DISTINCT staged source rows (<= 256, ~200 in patch order) are
// multiplied by the 32 x (S*NO) weight block ONCE (two M = 128 blocks, K = 32: 64 KB of TMEM stores per tile), the
// projected rows are parked in shared memory and every output row sums its nine NO-vectors from there.  The fp32-FMA
// kernel this replaces (narrow_conv.cuh, narrow_out_fwd_kernel) holds 108 weights per thread and runs 8 warps per SM.
// Precision: error-compensated 3xTF32 as every tcgen05 kernel here (P = A_hi [W_hi; W_lo] + A_lo W_hi, fp32 accumulation),
// the nine-term sums in fp32 in slot order (deterministic).
//
// Tile plan: the FORWARD tile plan of the layer's spiral table (tables.tile_plan: plan_cnt / plan_src / plan_cell; the
// cell word of (slot s, tile row r) holds the staged position p of src(r, s) as p*128 + 64*(p & 1)), rcap <= 256.
// Warp roles (24 warps): 0..7 epilogue, two groups of four warps that take alternate tiles (TMEM -> P tile in shared
// memory -> gather-sum -> coalesced store; with one group the epilogue bounded the kernel: 1.00 ms at 1024 meshes) |
// 8..19 splitters: chunk g = 2*tile + block goes to set g % 3, warp % 4 = TMEM lane quarter, staged rows -> hi/lo ->
// tcgen05.st.16x256b |
// 20, 21 loaders (warp w takes the tiles it = w (mod 2) into stage it % NTS, NTS even: ONE producer per stage barrier,
// so its parity waits are never two phases away; the stage's barrier is armed by the copies themselves --
// cp.async.mbarrier.arrive -- so a loader never waits for data and has two tiles in flight) | 22 idle | 23 TMEM
// allocation + MMA issue.
// TMEM columns: [0, 128) two accumulator sets of two blocks x 32 columns (the three 3xTF32 terms are three N = 32 MMAs
// into the same columns); [128, 512) three A stages of two blocks x (32 hi + 32 lo) columns.
#pragma once
#include "spiral_conv_tile.cuh"

namespace sdvae {
namespace tile {

constexpr int kOEpiGroups = 2;                                        // epilogue groups of 4 warps, tiles alternate between them
constexpr int kOFirstSplitWarp = 4 * kOEpiGroups;                     // 8
constexpr int kOSplitSets = 3;                                        // chunk g = 2*tile + block goes to set g % 3
constexpr int kOFirstLoadWarp = kOFirstSplitWarp + 4 * kOSplitSets;   // 20
constexpr int kOLoadWarps = 3;
constexpr int kOMmaWarp = kOFirstLoadWarp + kOLoadWarps;              // 23
constexpr int kOThreads = (kOMmaWarp + 1) * 32;                       // 768 -> 80 registers per thread
constexpr int kOMaxStages = 4;                                        // tile-stage ring depth: 4 or 2 (two loader warps, warp w owns the stages = w mod 2)
constexpr int kOBlocks = 2;                                           // M = 128 blocks of staged rows per tile
constexpr int kOAStages = 3;                                          // TMEM A ring (128 columns each)
constexpr int kOAColBase = 128;
constexpr int kOMaxRcap = 128 * kOBlocks;
constexpr int kOPStride = 29;                                         // floats per projected row in shared memory (odd: conflict-free writes)
constexpr int kOPBytes = kOMaxRcap * kOPStride * 4;

struct OutArgs {
    const float* in;              // [B, in_rows, 32]
    const int* plan_cnt;          // [L]
    const int* plan_src;          // [L, rcap/2]
    const uint32_t* plan_cell;    // [L, S*128]
    const float* W;               // [NO, S*32] the layer's weight
    const float* bias;            // [NO] or nullptr
    float* out;                   // [B, out_rows, NO]
    int B, in_rows, out_rows, L, S, NO, rcap, nts;
};

struct OutCfg {
    static size_t stage_bytes(int S, int rcap) { return (size_t)rcap * 128 + (size_t)S * 512; }
    static int stages(int S, int rcap) {
        const long long budget = 227LL * 1024 - 1024 - 512 - kTBChunk - 2 * kOPBytes;
        long long st = budget / (long long)stage_bytes(S, rcap);
        return st >= 4 ? 4 : (st >= 2 ? 2 : 0);
    }
    static size_t smem_bytes(int S, int rcap, int nts) { return 1024 + kTBChunk + (size_t)nts * stage_bytes(S, rcap) + 2 * kOPBytes + 512; }
};

#ifndef SDVAE_ABL_OUT
#define SDVAE_ABL_OUT 0            // compile-time ablation mask of tuning builds: 1 no split / TMEM stores, 2 no MMAs, 4 no epilogue work
#endif

__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void epi_bar_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
// arrive on `bar` (no pending-count increment: the barrier's expected count includes this arrival) once every cp.async
// this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// ST, NOT: compile-time S and NO (9, 3: the loops of the epilogue unroll, all shared-memory reads of a row are in
// flight at once), or 0, 0 for run-time values.
template <int ST, int NOT>
__global__ void __launch_bounds__(kOThreads, 1)
pt_kernel(const OutArgs a) {
    const int S = ST > 0 ? ST : a.S, NO = NOT > 0 ? NOT : a.NO, J = S * NO;
    const int NTS = a.nts;
    const int ROWS_BYTES = a.rcap * 128;
    const int STAGE_BYTES = ROWS_BYTES + S * 512;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* B_s = smem;                                    // [64][128 B] weight image: rows j < 32 hi, 32 + j lo of Wd[j = s*NO + n][c]
    uint8_t* T_s = B_s + kTBChunk;                          // [NTS] tile stages: rows | cell words
    float* P_s0 = reinterpret_cast<float*>(T_s + (size_t)NTS * STAGE_BYTES);  // [2][256][29] projected staged rows of the tile (double-buffered:
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(P_s0) + 2 * kOPBytes);   // one CTA barrier per tile)
    uint64_t* tile_full = bars;                             // [NTS]  loader lanes (32 asynchronous arrivals) -> splitters, epilogue
    uint64_t* tile_empty = bars + kOMaxStages;              // [NTS]  8 splitter warps + 4 epilogue warps -> loader
    uint64_t* a_full = bars + 2 * kOMaxStages;              // [3]    splitters (8 warps) -> MMA
    uint64_t* a_empty = a_full + kOAStages;                 // [3]    MMA (commit) -> splitters
    uint64_t* t_full = a_empty + kOAStages;                 // [2]    MMA (commit) -> epilogue
    uint64_t* t_empty = t_full + 2;                         // [2]    epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kOMaxStages; ++i) { mbar_init(tile_full + i, 32); mbar_init(tile_empty + i, 4 * kOBlocks + kTEpilogueWarps); }
        for (int i = 0; i < kOAStages; ++i) { mbar_init(a_full + i, 4 * kOBlocks); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, kTEpilogueWarps); }
        fence_barrier_init();
    }
    if (warp == kOMmaWarp) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    // weight image of the projection, split and swizzled in place (no pack launch): image row j = s*NO + n
    for (int t = tid; t < 64 * 32; t += kOThreads) {
        const int kk = t & 31, j = t >> 5, n = j & 31, part = j >> 5;
        float w = 0.f;
        if (n < J) w = __ldg(a.W + (size_t)(n % NO) * S * 32 + (n / NO) * 32 + kperm(kk));
        float hi, lo;
        split_tf32f(w, hi, lo);
        *reinterpret_cast<float*>(B_s + sw128_off(j, kk >> 2) + (kk & 3) * 4) = part ? lo : hi;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = a.B * a.L;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;

    if (warp == kOMmaWarp) {
        // ================= MMA issuer: per tile one wait, 2 blocks x 8 MMAs, two commits =================
        if (elect_one()) {
            constexpr uint32_t IDESC = idesc_tf32(kBM, kTNT);
            const uint64_t desc_hi = smem_desc_sw128(smem_u32(B_s));          // image rows 0..31: W_hi
            const uint64_t desc_lo = desc_hi + (uint64_t)((32 * 128) >> 4);    //            32..63: W_lo
            int sa = 0; uint32_t aph = 0;
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                const int sd = it & 1;
                mbar_wait(t_empty + sd, (uint32_t)(((it >> 1) & 1) ^ 1));
                mbar_wait(a_full + sa, aph);
                tc_fence_after();
                if (!(SDVAE_ABL_OUT & 2))
#pragma unroll
                for (int blk = 0; blk < kOBlocks; ++blk) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(sd * 64 + blk * 32);
                    const uint32_t a_hi = tmem_base + (uint32_t)(kOAColBase + sa * 128 + blk * 64), a_lo = a_hi + 32;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_tf32_ts(d_tmem, a_hi + k * 8, desc_hi + (uint64_t)(2 * k), IDESC, k != 0);
                        umma_tf32_ts(d_tmem, a_hi + k * 8, desc_lo + (uint64_t)(2 * k), IDESC, 1u);
                        umma_tf32_ts(d_tmem, a_lo + k * 8, desc_hi + (uint64_t)(2 * k), IDESC, 1u);
                    }
                }
                umma_commit(a_empty + sa);
                umma_commit(t_full + sd);
                if (++sa == kOAStages) { sa = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= kOFirstLoadWarp) {
        // ================= loaders: warp lw < 2 takes the tiles it = lw (mod 2); stage it % NTS (NTS even) =================
        const int lw = warp - kOFirstLoadWarp;
        const int q = lane & 7, rsub = lane >> 3;
        const int n_cell16 = (S * 512) >> 4;
        constexpr int PV = kOMaxRcap / 32;
        long long t0 = (long long)blockIdx.x + (long long)lw * gridDim.x;
        int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
#pragma unroll 1
        for (int it = lw; lw < 2 && it < my_tiles; it += 2) {
            const int ts = it % NTS;
            const uint32_t tph = (uint32_t)((it / NTS) & 1);
            const uint32_t stage_a = smem_u32(T_s) + (uint32_t)ts * (uint32_t)STAGE_BYTES;
            // odd positions: high half first (the parity of staged row e = 32j + 4t + rsub is that of rsub)
            const uint32_t dst_rows = stage_a + (uint32_t)rsub * 128u + (((uint32_t)q * 16u) ^ ((uint32_t)(rsub & 1) << 6));
            const uint32_t dst_cell = stage_a + (uint32_t)ROWS_BYTES;
            PlanRegs<PV> now;
            plan_fetch(now, a.plan_cnt, a.plan_src, jt, 1, 0, a.rcap, rsub);
            const char* gb = reinterpret_cast<const char*>(a.in + (size_t)b * a.in_rows * 32 + 4 * q);
            const char* cell_g = reinterpret_cast<const char*>(a.plan_cell + (size_t)jt * S * 128);
            mbar_wait_relaxed(tile_empty + ts, tph ^ 1);
#pragma unroll
            for (int j = 0; j < PV; ++j) {
                if (32 * j < now.n) {
                    const uint32_t w[4] = {now.w[j].x, now.w[j].y, now.w[j].z, now.w[j].w};
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const uint32_t row = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                                     ::"r"(dst_rows + (uint32_t)(32 * j + 4 * t) * 128u), "l"(gb + (size_t)row * 128u));
                    }
                }
            }
#pragma unroll 1
            for (int i = lane; i < n_cell16; i += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_cell + (uint32_t)i * 16u), "l"(cell_g + i * 16));
            cp_async_arrive_noinc(smem_u32(tile_full + ts));   // fires when this lane's copies have landed
            for (int k = 0; k < 2; ++k) {
                b += db; jt += djt;
                if (jt >= a.L) { jt -= a.L; ++b; }
            }
        }
    } else if (warp < kOFirstSplitWarp) {
        // ================= epilogue: projected rows TMEM -> shared memory, nine-term sums, coalesced store =================
        const int grp = warp >> 2;                             // tiles it = grp (mod 2)
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;                          // TMEM lane / tile row of this thread
        const uint32_t cell_idx = (uint32_t)((r >> 5) * 32 + (r & 7) * 4 + ((r >> 3) & 3)) * 4u;
        long long t0 = (long long)blockIdx.x + (long long)grp * gridDim.x;
        int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
        float bias[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.bias != nullptr)
            for (int n = 0; n < NO && n < 4; ++n) bias[n] = __ldg(a.bias + n);
#pragma unroll 1
        for (int it = grp; it < my_tiles; it += kOEpiGroups) {
            const int st = it & 1;
            const int ts = it % NTS;
            // the staged positions of this row's S source rows first: the tile stage is released before the MMAs of the
            // tile have even been issued (mbarrier.arrive is a release: the reads above it are done)
            mbar_wait_a<64>(smem_u32(tile_full + ts), (uint32_t)((it / NTS) & 1));
            const uint32_t cbase = smem_u32(T_s) + (uint32_t)ts * (uint32_t)STAGE_BYTES + (uint32_t)ROWS_BYTES + cell_idx;
            uint32_t pw[9];
#pragma unroll
            for (int s = 0; s < 9; ++s)
                pw[s] = s < S ? ((lds32u(cbase + (uint32_t)s * 512u) & 0xffffu) >> 7) * (uint32_t)kOPStride + (uint32_t)(s * NO) : 0u;
            warp_arrive(tile_empty + ts, lane);
            mbar_wait_relaxed(t_full + st, (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            float* P_s = P_s0 + (size_t)st * (kOPBytes / 4);
            if (!(SDVAE_ABL_OUT & 4))
#pragma unroll
            for (int blk = 0; blk < kOBlocks; ++blk) {
                const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(st * 64 + blk * 32);
                float* prow = P_s + (size_t)(blk * 128 + r) * kOPStride;
                float v0[16], v1[16];
                tmem_ld16(t_row, v0);
                tmem_ld16(t_row + 16, v1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) prow[j] = v0[j];
#pragma unroll
                for (int j = 0; j < kOPStride - 16; ++j) prow[16 + j] = v1[j];
            }
            tc_fence_before();
            warp_arrive(t_empty + st, lane);
            epi_bar_sync(1 + grp);                             // the whole P tile is in shared memory
            float acc[4] = {bias[0], bias[1], bias[2], bias[3]};
            if (!(SDVAE_ABL_OUT & 4)) {
                float pv[9][4];
#pragma unroll
                for (int s = 0; s < 9; ++s)
#pragma unroll
                    for (int n = 0; n < 4; ++n) pv[s][n] = (s < S && n < NO) ? P_s[pw[s] + n] : 0.f;
#pragma unroll
                for (int s = 0; s < 9; ++s)
#pragma unroll
                    for (int n = 0; n < 4; ++n) acc[n] += pv[s][n];
            }
            const int row = jt * kBM + r;
            if (row < a.out_rows) {
                float* o = a.out + ((size_t)b * a.out_rows + row) * NO;
#pragma unroll
                for (int n = 0; n < 4; ++n)
                    if (n < NO) o[n] = acc[n];
            }
            epi_bar_sync(1 + grp);                             // every warp of the group is done with its P buffer
            for (int k = 0; k < kOEpiGroups; ++k) {
                b += db; jt += djt;
                if (jt >= a.L) { jt -= a.L; ++b; }
            }
        }
    } else {
        // ================= splitters: staged rows 128*blk + 32*q4 .. -> hi / lo -> TMEM A stage =================
        const int set = (warp - kOFirstSplitWarp) >> 2;
        const int q4 = warp & 3;
        const int l4 = lane >> 2, qq = lane & 3;
        const uint32_t T_a = smem_u32(T_s);
        const uint32_t t_q = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)kOAColBase;
        const uint32_t bar_tile_full = smem_u32(tile_full), bar_tile_empty = smem_u32(tile_empty);
        const uint32_t bar_a_full = smem_u32(a_full), bar_a_empty = smem_u32(a_empty);
        // byte offsets of this thread's four staged rows in block 0 / block 1 (positions p = 128 blk + 32 q4 + 8 k + l4;
        // rows past the stage's capacity shadow its last row: their projections are never gathered)
        uint32_t off[2][4];
#pragma unroll
        for (int bk = 0; bk < 2; ++bk)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int p = bk * 128 + q4 * 32 + 8 * k + l4;
                p = p < a.rcap ? p : a.rcap - 1;
                off[bk][k] = (uint32_t)(p * 128 + 64 * (p & 1)) + (uint32_t)qq * 16u;
            }
#pragma unroll 1
        for (int g = set; g < 2 * my_tiles; g += kOSplitSets) {
            const int it = g >> 1, blk = g & 1;
            const int st = it % kOAStages;
            const uint32_t ph = (uint32_t)((it / kOAStages) & 1);
            const int ts = it % NTS;
            const uint32_t stage_a = T_a + (uint32_t)ts * (uint32_t)STAGE_BYTES;
            const uint32_t t_lane = t_q + (uint32_t)(blk * 64);
            mbar_wait_a<64>(bar_tile_full + (uint32_t)ts * 8u, (uint32_t)((it / NTS) & 1));
            float4 X[4], Y[4];
            if (!(SDVAE_ABL_OUT & 1))
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ax = stage_a + (blk ? off[1][k] : off[0][k]);
                X[k] = lds128(ax);
                Y[k] = lds128(ax ^ 64u);
            }
            mbar_wait_a<32>(bar_a_empty + (uint32_t)st * 8u, ph ^ 1);
            __syncwarp();
            tc_fence_after();
            if (!(SDVAE_ABL_OUT & 1))
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
                float r[32];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 P = X[gg * 2 + h];
                    const float4 Q = Y[gg * 2 + h];
                    const float v8[8] = {P.x, P.y, P.z, P.w, Q.x, Q.y, Q.z, Q.w};
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const float2 v2 = make_float2(v8[2 * n], v8[2 * n + 1]);
                        const float2 hi = make_float2(__uint_as_float(__float_as_uint(v2.x) & 0xffffe000u),
                                                      __uint_as_float(__float_as_uint(v2.y) & 0xffffe000u));
                        const float2 lo = sub2(v2, hi);
                        r[4 * n + 2 * h] = hi.x; r[4 * n + 2 * h + 1] = hi.y;
                        r[16 + 4 * n + 2 * h] = lo.x; r[16 + 4 * n + 2 * h + 1] = lo.y;
                    }
                }
                tmem_st_16x256b_x8(t_lane + (uint32_t)(st * 128) + ((uint32_t)(16 * gg) << 16), r);
            }
            tmem_st_wait();
            tc_fence_before();
            if (lane == 0) {
                mbar_arrive_a(bar_a_full + (uint32_t)st * 8u);
                mbar_arrive_a(bar_tile_empty + (uint32_t)ts * 8u);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kOMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace tile
}  // namespace sdvae
