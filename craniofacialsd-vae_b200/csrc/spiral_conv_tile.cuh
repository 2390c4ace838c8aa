// Spiral convolution on tcgen05 with TILE-LOCAL STAGING and stage-granular hand-offs (32 channels per slot,
// 32 output channels): forward (model.py:27-41 + F.elu, model.py:68,84) and backward-to-input (autograd of
// model.py:34,40 = the same contraction over the inverse table, a deterministic in-order scatter-add).
//
//   y[m, n] = epi( sum_{s,c} A[m, s*32 + c] * W[n, s*32 + c] ),   A[m, s*32 + c] = sum_{r in cell(m, s)} x[r, c]
//
// Why this kernel exists (profiles/r01_*, r02_*): gc_umma_kernel copies 9 x 128 gathered rows per 128-row tile
// through the L2->SM fabric (the fabric, ~7 TB/s, bounded it, not HBM) and hands every 32-wide K chunk from role
// to role through its own mbarrier round (the MMA-issuing thread spent 350 clk per chunk on waits and commits
// against 192 clk of tensor-pipe time).  Here
//   * a tile's DISTINCT source rows (~200 of the 1152 gathered, once the level is numbered patch-wise) are
//     copied ONCE into a tile stage together with the tile's cell words (host-built TILE PLAN, tables.tile_plan);
//   * the A operand is handed to the MMA thread per STAGE of three K chunks (a third of the barrier rounds,
//     24 MMAs per wait + commit);
//   * the splitters gather from shared memory (almost) without bank conflicts for arbitrary rows: they write the
//     TMEM A operand with tcgen05.st.16x256b (four threads per row), the four threads of a row read one 64-byte
//     half of it with one LDS.128; staged rows at odd positions of the tile stage are stored high half first, and
//     the plan builder places the rows (max-cut) so that the two rows of an 8-lane phase mostly sit at positions
//     of different parity (~75 % of the phases conflict-free, the rest 2-way).  [Reading opposite halves by lane
//     parity instead is always conflict-free but costs 32 selects per unit and measured 0.2 ms slower,
//     profiles/r02_tile_kernel.md.]  The price is a fixed permutation of the 32 channels of a K chunk
//     (kperm), applied to the weight image by the packer.
// Precision, weight image, MMA form (TS: A from TMEM, B = resident weight image), epilogue: as gc_umma_kernel
// (error-compensated 3xTF32, fp32 accumulation in TMEM).
//
// Tile plan (per tile t of 128 output rows; all tables static, built once on the host):
//   plan_cnt [L]             distinct source rows of the tile (<= rcap)
//   plan_src [L, rcap/2]     those rows, 16-bit pairs in loader-lane order (umma::plan_fetch, S = 1)
//   plan_cell[L, S*128]      one word per (slot s, tile row r), stored at  s*128 + (r>>5)*32 + (r&7)*4 + ((r>>3)&3)
//                            (the four words of a splitter thread are one 16-byte load):
//                              bits 0..15 byte offset of the LOW 64-byte half of the cell's first row in the tile stage
//                              (position p: p*128 + 64*(p & 1) -- rows at odd positions are stored high half first),
//                              bits 16..20 rows in the cell, bits 21.. offset of its 2nd, 3rd, ... rows in plan_ext
//   plan_ext [L, ecap]       16-bit byte offsets (same form) of the 2nd, 3rd ... rows of the cells (RAGGED plans only)
#pragma once
#include "spiral_conv_umma.cuh"

namespace sdvae {
namespace tile {

using namespace umma;

constexpr int kTEpilogueWarps = 4;
constexpr int kTMaxStages = 3;                                  // tile-stage ring depth limit (one loader warp each)
constexpr int kTFirstSplitWarp = kTEpilogueWarps;               // 4
// Warp layout for NSETS splitter sets (4 warps each, one per TMEM lane quarter):
//   warps 0..3 epilogue | 4 .. 4+4*NSETS-1 splitters | then kTMaxStages loader warps | then the MMA warp.
// The splitters are LATENCY-bound (a unit is a serial chain of two barrier waits, two dependent shared-memory
// round trips, ~70 ALU instructions, two TMEM stores and their completion wait: ~1300 clk with the SM's other
// warps competing), so throughput = sets in flight / unit latency.  Measured (profiles/r02_tile_kernel.md): 4 sets at
// 88 registers beat 5 at 80 and 6 at 64 (2.04 / 2.13 / 2.24 ms); launch_tile uses 4.
template <int NSETS>
struct TileWarps {
    static constexpr int kSplitWarps = 4 * NSETS;
    static constexpr int kFirstLoadWarp = kTFirstSplitWarp + kSplitWarps;
    static constexpr int kMmaWarp = kFirstLoadWarp + kTMaxStages;
    static constexpr int kThreads = (kMmaWarp + 1) * 32;        // 4 sets: 768, 5: 896, 6: 1024
    // setmaxnreg budgets per warpgroup (sum = kThreads * registers at launch)
    static constexpr int kRegsEpilogue = NSETS == 4 ? 64 : (NSETS == 5 ? 56 : 64);
    static constexpr int kRegsSplit = NSETS == 4 ? 88 : (NSETS == 5 ? 80 : 64);
    static constexpr int kRegsLoad = NSETS == 4 ? 64 : (NSETS == 5 ? 48 : 64);
    static constexpr bool kRealloc = NSETS != 6;
};
constexpr int kTMaxRcap = 288;                                  // distinct rows per tile (multiple of 32)
constexpr int kTChunksPerStage = 3;                             // K chunks per TMEM A stage
constexpr int kTNT = 32;
constexpr int kTAccCols = 4 * kTNT;                             // two accumulators of 2*NT columns
constexpr int kTStageCols = kTChunksPerStage * 64;              // 192: (32 hi + 32 lo) per chunk
constexpr int kTBChunk = 2 * kTNT * 128;                        // weight image bytes per K chunk

struct TileArgs {
    const float* in;              // [B, in_rows, 32]
    const int* plan_cnt;          // [L]
    const int* plan_src;          // [L, rcap/2]
    const uint32_t* plan_cell;    // [L, S*128]
    const uint16_t* plan_ext;     // [L, ecap] (RAGGED) or nullptr
    const float* wimg;            // packed weight image with kperm applied (sdvae_tc_pack_weights, perm = 1)
    const float* bias;            // [32] or nullptr
    const float* gate;            // EPI_GATE: out *= elu'(gate), aligned with out
    float* out;                   // [B, out_rows, ldo]
    int B, in_rows, out_rows, L, S, rcap, ecap;
    int ldo, epi, nts;
    int dbg;
};

struct TileCfg {
    static size_t b_bytes(int S) { return (size_t)S * kTBChunk; }
    static size_t stage_bytes(int S, int rcap, int ecap) { return (size_t)rcap * 128 + (size_t)S * 512 + (size_t)ecap * 2; }
    static int stages(int S, int rcap, int ecap) {
        const long long budget = 227LL * 1024 - 1024 /*align*/ - 512 /*barriers*/ - (long long)b_bytes(S) - kOutStageBytes;
        long long st = budget / (long long)stage_bytes(S, rcap, ecap);
        return (int)(st > kTMaxStages ? kTMaxStages : st);
    }
    static size_t smem_bytes(int S, int rcap, int ecap, int nts) {
        return 1024 + b_bytes(S) + (size_t)nts * stage_bytes(S, rcap, ecap) + kOutStageBytes + 512;
    }
};

// 16 lanes x 64 columns: thread t holds lanes (t >> 2) and (t >> 2) + 8; register 4n + 2h + b = lane (t >> 2) + 8h,
// column 8n + 2(t & 3) + b   (cute SM100_TMEM_STORE_16dp256b8x)
__device__ __forceinline__ void tmem_st_16x256b_x8(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr),
          "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
          "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
          "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}

// ---- shared-memory access by 32-bit shared address (no generic-address arithmetic in the per-unit loops) ----
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16u(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void warp_arrive_a(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive_a(bar);
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
    return ok != 0;
}
// bounded wait with a short back-off between polls (a failed try_wait returns after ~40 clk whatever the hint says;
// 16 splitter warps polling back to back took a third of the SM's issue slots)
template <unsigned NS>
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_a(bar, parity)) return;
    int spins = 0;
    do {
        __nanosleep(NS);
        if (++spins > kSpinLimit) __trap();
    } while (!mbar_try_wait_a(bar, parity));
}

// ---- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): the kernel is bound by instruction issue ----
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
// elu_fast (common.cuh) on a pair: same polynomial / ex2 branches, the polynomial in FFMA2
__device__ __forceinline__ float2 elu_fast2(float2 v) {
    const float2 vv = v;
    float2 p = make_float2(2.7557319e-6f, 2.7557319e-6f);
    p = fma2(p, vv, make_float2(2.4801587e-5f, 2.4801587e-5f));
    p = fma2(p, vv, make_float2(1.9841270e-4f, 1.9841270e-4f));
    p = fma2(p, vv, make_float2(1.3888889e-3f, 1.3888889e-3f));
    p = fma2(p, vv, make_float2(8.3333333e-3f, 8.3333333e-3f));
    p = fma2(p, vv, make_float2(4.1666667e-2f, 4.1666667e-2f));
    p = fma2(p, vv, make_float2(1.6666667e-1f, 1.6666667e-1f));
    p = fma2(p, vv, make_float2(0.5f, 0.5f));
    p = fma2(p, vv, make_float2(1.0f, 1.0f));
    p = mul2(p, vv);
    const float2 e = add2(make_float2(__expf(v.x), __expf(v.y)), make_float2(-1.0f, -1.0f));
    const float nx = v.x > -0.5f ? p.x : e.x, ny = v.y > -0.5f ? p.y : e.y;
    return make_float2(v.x > 0.f ? v.x : nx, v.y > 0.f ? v.y : ny);
}

template <bool RAGGED, int NSETS>
__global__ void __launch_bounds__(TileWarps<NSETS>::kThreads, 1)
gt_kernel(const TileArgs a) {
    using TW = TileWarps<NSETS>;
    constexpr int kTSplitSets = NSETS, kTSplitWarps = TW::kSplitWarps, kTFirstLoadWarp = TW::kFirstLoadWarp;
    constexpr int kTMmaWarp = TW::kMmaWarp, kTThreads = TW::kThreads;
    const int S = a.S;
    const int NCH = S;                                      // one 32-wide K chunk per slot
    const int NTS = a.nts;
    const int ROWS_BYTES = a.rcap * 128;
    const int CELL_BYTES = S * 512;
    const int STAGE_BYTES = ROWS_BYTES + CELL_BYTES + a.ecap * 2;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* B_s = smem;                                    // [S][64][128 B]  resident weight image
    uint8_t* T_s = B_s + (size_t)NCH * kTBChunk;            // [NTS] tile stages: rows | cell words | ext
    uint8_t* O_s = T_s + (size_t)NTS * STAGE_BYTES;         // [128][144 B] output staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(O_s + kOutStageBytes);
    uint64_t* tile_full = bars;                             // [NTS]  loader    -> splitters
    uint64_t* tile_empty = bars + kTMaxStages;              // [NTS]  splitters -> loader
    uint64_t* a_full = bars + 2 * kTMaxStages;              // [2]    splitters -> MMA   (per A stage of 3 chunks)
    uint64_t* a_empty = a_full + 2;                         // [2]    MMA (commit) -> splitters
    uint64_t* t_full = a_empty + 2;                         // [2]    MMA (commit) -> epilogue
    uint64_t* t_empty = t_full + 2;                         // [2]    epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kTMaxStages; ++i) { mbar_init(tile_full + i, 1); mbar_init(tile_empty + i, kTSplitWarps); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + i, 4 * kTChunksPerStage); mbar_init(a_empty + i, 1);
            mbar_init(t_full + i, 1); mbar_init(t_empty + i, kTEpilogueWarps);
        }
        fence_barrier_init();
    }
    if (warp == kTMmaWarp) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    {
        const int n16 = NCH * kTBChunk / 16;
        const float4* src = reinterpret_cast<const float4*>(a.wimg);
        float4* dst = reinterpret_cast<float4*>(B_s);
#pragma unroll 1
        for (int i = tid; i < n16; i += kTThreads) dst[i] = __ldg(src + i);
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = a.B * a.L;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int G = my_tiles * NCH;                           // chunks this CTA processes
    const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;

    if (warp >= kTFirstLoadWarp) {
        if (TW::kRealloc) reg_dec<TW::kRegsLoad>();
        if (warp == kTMmaWarp) {
            // ================= MMA issuer: one wait + 24 MMAs + one commit per stage of three chunks =================
            if (elect_one()) {
                constexpr uint32_t IDESC1 = idesc_tf32(kBM, 2 * kTNT);
                constexpr uint32_t IDESC2 = idesc_tf32(kBM, kTNT);
                const uint64_t desc0 = smem_desc_sw128(smem_u32(B_s));
                const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
                int st = 0; uint32_t sph = 0;
                const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0;
                WaitClock w_afull(prof), w_tempty(prof);
                const long long t_begin = prof ? clock64() : 0;
#pragma unroll 1
                for (int it = 0; it < my_tiles; ++it) {
                    const int acc = it & 1;
                    w_tempty.timed([&] { mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1); });
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * kTNT);
#pragma unroll 1
                    for (int c0 = 0; c0 < NCH; c0 += kTChunksPerStage) {
                        w_afull.timed([&] { mbar_wait(a_full + st, sph); });
                        tc_fence_after();
                        if (!SDVAE_DBG_ON(a, 4))
#pragma unroll
                        for (int c = 0; c < kTChunksPerStage; ++c) {
                            const uint32_t a_hi = tmem_base + (uint32_t)(kTAccCols + st * kTStageCols + c * 64), a_lo = a_hi + 32;
                            const uint32_t dl = desc_lo0 + (uint32_t)((c0 + c) * (kTBChunk >> 4));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(dl + 2u * k);
                                umma_tf32_ts(d_tmem, a_hi + k * 8, bd, IDESC1, (c0 | c | k) != 0);
                                umma_tf32_ts(d_tmem, a_lo + k * 8, bd, IDESC2, 1u);
                            }
                        }
                        umma_commit(a_empty + st);
                        if (c0 + kTChunksPerStage >= NCH) umma_commit(t_full + acc);
                        if (++st == 2) { st = 0; sph ^= 1; }
                    }
                }
                if (prof) { g_prof[0] = clock64() - t_begin; g_prof[1] = w_afull.acc; g_prof[2] = w_tempty.acc; g_prof[3] = G; }
            }
            __syncwarp();
        } else {
            // ================= loaders: warp lw owns tile stage lw, one whole tile per pass =================
            const int lw = warp - kTFirstLoadWarp;
            if (lw < NTS) {
                const int q = lane & 7, rsub = lane >> 3;
                uint8_t* stage = T_s + (size_t)lw * STAGE_BYTES;
                const uint32_t dst_rows = smem_u32(stage) + (uint32_t)rsub * 128u + (((uint32_t)q * 16u) ^ ((uint32_t)(rsub & 1) << 6));   // odd positions: high half first
                const uint32_t dst_cell = smem_u32(stage) + (uint32_t)ROWS_BYTES;
                const int n_cell16 = (CELL_BYTES + a.ecap * 2) >> 4;       // cell words and ext are contiguous in the stage
                constexpr int PV = kTMaxRcap / 32;
                long long t0 = (long long)blockIdx.x + (long long)lw * gridDim.x;
                int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
                uint32_t tph = 0;
                const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0 && lw == 0;
                WaitClock w_tempty(prof), w_copy(prof);
                const long long t_begin = prof ? clock64() : 0;
#pragma unroll 1
                for (int it = lw; it < my_tiles; it += NTS) {
                    PlanRegs<PV> now;
                    plan_fetch(now, a.plan_cnt, a.plan_src, jt, 1, 0, a.rcap, rsub);
                    const float* base = a.in + (size_t)b * a.in_rows * 32 + 4 * q;
                    const char* cell_g = reinterpret_cast<const char*>(a.plan_cell + (size_t)jt * S * 128);
                    const char* ext_g = RAGGED ? reinterpret_cast<const char*>(a.plan_ext + (size_t)jt * a.ecap) : nullptr;
                    w_tempty.timed([&] { mbar_wait_relaxed(tile_empty + lw, tph ^ 1); });
                    if (!SDVAE_DBG_ON(a, 1))
                    {   // rows: staged row e = 32*j + 4*t + rsub, piece q (no swizzle: the splitter reads are conflict-free by construction)
                        const char* gb = reinterpret_cast<const char*>(base);
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
                    }
#pragma unroll 1
                    for (int i = lane; i < n_cell16; i += 32) {
                        const int off = i * 16;
                        const char* src = (!RAGGED || off < CELL_BYTES) ? cell_g + off : ext_g + (off - CELL_BYTES);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_cell + (uint32_t)off), "l"(src));
                    }
                    cp_async_commit();
                    w_copy.timed([&] { cp_async_wait<0>(); });
                    warp_arrive(tile_full + lw, lane);
                    tph ^= 1;
                    for (int k = 0; k < NTS; ++k) {
                        b += db; jt += djt;
                        if (jt >= a.L) { jt -= a.L; ++b; }
                    }
                }
                if (prof && lane == 0) { g_prof[8] = clock64() - t_begin; g_prof[9] = w_tempty.acc; g_prof[10] = w_copy.acc; }
            }
        }
    } else if (warp < kTFirstSplitWarp) {
        if (TW::kRealloc) reg_dec<TW::kRegsEpilogue>();
        // ================= epilogue (as gc_umma_kernel's forward epilogue, rows staged for coalesced stores) ======
        const int q4 = warp & 3;
        const int EPI = a.epi;
        const int ldo = a.ldo;
        const bool has_bias = (EPI == EPI_BIAS || EPI == EPI_BIAS_ELU) && a.bias != nullptr;
        int b = (int)blockIdx.x / a.L, jt = (int)blockIdx.x - b * a.L;
        const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0 && warp == 0;
        WaitClock w_tfull(prof);
        const long long t_begin = prof ? clock64() : 0;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            w_tfull.timed([&] { mbar_wait_relaxed(t_full + acc, (it >> 1) & 1); });
            tc_fence_after();
            const int r = jt * kBM + q4 * 32 + lane;
            const size_t m = (size_t)b * a.out_rows + (r < a.out_rows ? r : 0);
            const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * 2 * kTNT);
#pragma unroll 1
            for (int c0 = 0; c0 < kTNT; c0 += 16) {
                float v[16], d2[16];
                tmem_ld16(t_row + c0, v);
                tmem_ld16(t_row + kTNT + c0, d2);
                tmem_ld_wait();
                if (c0 + 16 >= kTNT) {
                    tc_fence_before();
                    warp_arrive(t_empty + acc, lane);
                }
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float2 t = add2(make_float2(v[j], v[j + 1]), make_float2(d2[j], d2[j + 1]));
                    v[j] = t.x; v[j + 1] = t.y;
                }
                if (has_bias) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bv = ldg4(a.bias + c0 + j);
                        const float2 t0 = add2(make_float2(v[j], v[j + 1]), make_float2(bv.x, bv.y));
                        const float2 t1 = add2(make_float2(v[j + 2], v[j + 3]), make_float2(bv.z, bv.w));
                        v[j] = t0.x; v[j + 1] = t0.y; v[j + 2] = t1.x; v[j + 3] = t1.y;
                    }
                }
                if (EPI == EPI_BIAS_ELU) {
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float2 t = elu_fast2(make_float2(v[j], v[j + 1]));
                        v[j] = t.x; v[j + 1] = t.y;
                    }
                }
                if (EPI == EPI_GATE) {
                    const float* grow = a.gate + m * ldo + c0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 gt = ldg4(grow + j);
                        v[j] *= elu_grad_from_out(gt.x); v[j + 1] *= elu_grad_from_out(gt.y);
                        v[j + 2] *= elu_grad_from_out(gt.z); v[j + 3] *= elu_grad_from_out(gt.w);
                    }
                }
                float4* srow = reinterpret_cast<float4*>(O_s + (q4 * 32 + lane) * kOutRowBytes + c0 * 4);
#pragma unroll
                for (int j = 0; j < 16; j += 4) srow[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
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
            b += db; jt += djt;
            if (jt >= a.L) { jt -= a.L; ++b; }
        }
        if (prof && lane == 0) { g_prof[16] = clock64() - t_begin; g_prof[17] = w_tfull.acc; }
    } else {
        if (TW::kRealloc) reg_inc<TW::kRegsSplit>();
        // ================= splitters =================
        // set k takes the chunks g = k (mod 4) of the CTA's chunk sequence (tile iteration, slot); its four warps
        // own the four TMEM lane quarters.  Thread (l4 = lane >> 2, qq = lane & 3) serves the tile rows
        // 32*q4 + 16*g + 8*h + l4 (g, h in {0, 1}): for each it reads the two 16-byte pieces qq and qq + 4 of the
        // staged row(s) of the cell (the plan word holds the byte offset of the row's low half).
        // The kernel is bound by this loop -- above all by its TMEM stores and their completion waits
        // (profiles/r02_tile_kernel.md): everything per unit that is not a row read, a split or a TMEM store is
        // kept out of it.
        const int set = (warp - kTFirstSplitWarp) >> 2;
        const int q4 = warp & 3;
        const int l4 = lane >> 2, qq = lane & 3;
        const uint32_t offX = (uint32_t)qq * 16u;                          // low half (pieces 0..3); high half: ^ 64
        const uint32_t T_a = smem_u32(T_s);
        const uint32_t cell_off = (uint32_t)ROWS_BYTES + (uint32_t)(q4 * 32 + l4 * 4) * 4u;
        const uint32_t ext_off = (uint32_t)(ROWS_BYTES + CELL_BYTES);
        const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)kTAccCols;
        const uint32_t bar_tile_full = smem_u32(tile_full), bar_tile_empty = smem_u32(tile_empty);
        const uint32_t bar_a_full = smem_u32(a_full), bar_a_empty = smem_u32(a_empty);
        int ts = 0; uint32_t tph = 0;
        uint32_t stage_a = T_a;
        const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0 && warp == kTFirstSplitWarp;
        long long seg[6] = {0, 0, 0, 0, 0, 0};
        const long long t_begin = prof ? clock64() : 0;
        long long t_prev = t_begin;
#define SDVAE_SEG(i) do { if (prof) { const long long t_now_ = clock64(); seg[i] += t_now_ - t_prev; t_prev = t_now_; } } while (0)
        int first = set;                                         // first slot of this set in the current tile: chunk g = it*NCH + ch belongs to set g % NSETS
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait_a<64>(bar_tile_full + (uint32_t)ts * 8u, tph);
            SDVAE_SEG(0);
            const uint32_t baseX = stage_a + offX;
#pragma unroll 1
            for (int ch = first; ch < NCH; ch += kTSplitSets) {
                const int sg = it * (NCH / kTChunksPerStage) + ((ch * 43) >> 7);     // stage round of the chunk (ch / 3)
                const int sub = ch - 3 * ((ch * 43) >> 7);
                const int st = sg & 1;
                const uint32_t sph = (uint32_t)((sg >> 1) & 1);
                const uint4 cw = lds128u(stage_a + cell_off + (uint32_t)ch * 512u);
                const uint32_t words[4] = {cw.x, cw.y, cw.z, cw.w};
                const uint32_t t_a = t_lane + (uint32_t)(st * kTStageCols + sub * 64);
                // gather (cell words -> staged rows, in-order cell sums) BEFORE the TMEM stage is known to be free; the
                // hi/lo split and the TMEM stores follow the wait (holding the 64 split values across the wait instead
                // costs more registers than the kernel has)
                float4 X[4], Y[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t w = words[k];
                    const uint32_t ax = baseX + (w & 0xffffu);
                    X[k] = lds128(ax);
                    Y[k] = lds128(ax ^ 64u);
                    if (RAGGED) {
                        const int cnt = (int)((w >> 16) & 0x1fu);
                        if (cnt == 0) { X[k] = make_float4(0.f, 0.f, 0.f, 0.f); Y[k] = X[k]; }
                        uint32_t ea = stage_a + ext_off + ((w >> 21) << 1);
#pragma unroll 1
                        for (int e = 1; e < cnt; ++e, ea += 2) {   // in-order sum: deterministic scatter-add
                            const uint32_t ax2 = baseX + lds16u(ea);
                            const float4 X2 = lds128(ax2);
                            const float4 Y2 = lds128(ax2 ^ 64u);
                            X[k].x += X2.x; X[k].y += X2.y; X[k].z += X2.z; X[k].w += X2.w;
                            Y[k].x += Y2.x; Y[k].y += Y2.y; Y[k].z += Y2.z; Y[k].w += Y2.w;
                        }
                    }
                }
                SDVAE_SEG(1);
                mbar_wait_a<32>(bar_a_empty + (uint32_t)st * 8u, sph ^ 1);   // the MMAs of the stage's previous round are done
                __syncwarp();                                    // the wait (and the cell loops above) diverge; tcgen05.st is warp-collective
                tc_fence_after();
                SDVAE_SEG(3);
#pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                    float r[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 P = X[gg * 2 + h];          // piece qq     : channels 4qq .. 4qq+3   -> n = 0, 1
                        const float4 Q = Y[gg * 2 + h];          // piece qq + 4 : channels 16+4qq ..      -> n = 2, 3
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
                    if (!SDVAE_DBG_ON(a, 16)) tmem_st_16x256b_x8(t_a + ((uint32_t)(16 * gg) << 16), r);
                }
                tmem_st_wait();
                SDVAE_SEG(4);
                tc_fence_before();
                if (lane == 0) mbar_arrive_a(bar_a_full + (uint32_t)st * 8u);     // (wait::st is warp-collective: every lane is done)
                SDVAE_SEG(5);
            }
            if (lane == 0) mbar_arrive_a(bar_tile_empty + (uint32_t)ts * 8u);     // this warp is done with the tile stage
            stage_a += (uint32_t)STAGE_BYTES;
            if (++ts == NTS) { ts = 0; tph ^= 1; stage_a = T_a; }
            first -= NCH % kTSplitSets;
            if (first < 0) first += kTSplitSets;
            SDVAE_SEG(2);
        }
#undef SDVAE_SEG
        if (prof && lane == 0) { g_prof[24] = clock64() - t_begin; for (int i = 0; i < 6; ++i) g_prof[25 + i] = seg[i]; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace tile
}  // namespace sdvae
