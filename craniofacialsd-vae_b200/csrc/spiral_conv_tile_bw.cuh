// SpiralConv weight gradient on tcgen05 with TILE-LOCAL STAGING (32 channels per slot, <= 32 output channels per pass).
// Reference: autograd of nn.Linear in model.py:40 over the (never materialised) gather of model.py:34.
//
//   dW[n, s*32 + c] = sum_m g[m, n] * x[src(m, s), c]          db[n] = sum_m g[m, n]
//
// The contraction is bw_umma_kernel's (spiral_conv_umma_bw.cuh): D[(s,c), n] += A^T[(s,c), m] * g[m, n] with the mesh
// rows as the K dimension, M = 128 = four slots x 32 channels per accumulator block, a row of ones behind the last
// slot for db, g as the MN-major shared-memory operand (hi image / lo image), 3xTF32 as three N = 32 MMAs per k-step,
// accumulators drained every `flush` tiles into per-CTA partials (the tensor core adds with truncation).
// What changes (profiles/r02_*): bw_umma_kernel copies 9 x 128 gathered rows per tile through the L2->SM fabric,
// which bounds it (~7 TB/s); here the tile's DISTINCT source rows are copied once into a tile stage together
// with the tile's cell words (the forward TILE PLAN of spiral_conv_tile.cuh, tables.tile_plan) and the transposing
// splitters read them there: thread (slot s, channel c) reads word c of the staged row of tile row m -- all 32
// lanes of a warp read the same 128-byte row, so arbitrary rows cost no bank conflicts.  Sixteen splitter warps
// (four sets) instead of eight, one arrival per warp on every barrier, one elected MMA thread that polls the next
// chunk's barrier in the middle of the current chunk's MMAs (tools/mma_loop_bench.cu).
//
// Warp roles (24 warps): 0..3 g staging + accumulator drain | 4..19 splitters, set k = chunks c = k (mod 4) of the
// CTA's chunk sequence (tile, block, 32-row group), warp % 4 = TMEM lane quarter | 20..22 loaders (one tile stage
// each) | 23 TMEM allocation + MMA issue.
// TMEM columns: [0, 192) two accumulator sets of NBLK (<= 3) x 32 columns; [192, 512) A ring: five stages of one chunk
// (32 hi + 32 lo columns).
#pragma once
#include "spiral_conv_tile.cuh"
#include "spiral_conv_umma_bw.cuh"

namespace sdvae {
namespace tile {

constexpr int kWFirstSplitWarp = 4;
constexpr int kWSplitSets = 4;
constexpr int kWFirstLoadWarp = kWFirstSplitWarp + 4 * kWSplitSets;     // 20
constexpr int kWMmaWarp = kWFirstLoadWarp + kTMaxStages;                // 23
constexpr int kWThreads = (kWMmaWarp + 1) * 32;                         // 768 -> 80 registers per thread
constexpr int kWAStages = 5;                                            // TMEM A ring (64 columns each)
constexpr int kWAColBase = 192;
constexpr int kWMaxBlocks = 3;                                          // accumulator blocks (S*32 + 1 <= 384 M rows)

struct TileBwArgs {
    const float* in;              // [B, in_rows, in_ld], the 32 channels of this pass start at `in`
    const int* plan_cnt;          // forward tile plan of the layer's table (tables.tile_plan): [L]
    const int* plan_src;          //   [L, rcap/2]
    const uint32_t* plan_cell;    //   [L, S*128]
    const float* g;               // [B, out_rows, g_ld], n_real columns from `g`
    float* part;                  // per-CTA partial dW (layout of umma::BwUmmaArgs::part), zero-initialised by the caller
    float* part_b;                // per-CTA partial db, or nullptr
    int in_ld, g_ld, part_ld, part_cta, partb_cta;
    int B, in_rows, out_rows, L, S, rcap, n_real, nts;
    int flush;                    // tiles per accumulator drain (>= 1)
};

struct TileBwCfg {
    static size_t stage_bytes(int S, int rcap) { return (size_t)rcap * 128 + (size_t)S * 512; }
    static size_t acc_bytes(int S) { return (size_t)((S * 32 + 1 + 127) >> 7) * 32 * 128 * 4; }     // drain accumulators
    static int stages(int S, int rcap) {
        const long long budget = 227LL * 1024 - 1024 - 512 - 2LL * umma::kGStage - (long long)acc_bytes(S);
        long long st = budget / (long long)stage_bytes(S, rcap);
        return (int)(st > kTMaxStages ? kTMaxStages : st);
    }
    static size_t smem_bytes(int S, int rcap, int nts) {
        return 1024 + 2 * (size_t)umma::kGStage + (size_t)nts * stage_bytes(S, rcap) + acc_bytes(S) + 512;
    }
};

#ifndef SDVAE_ABL
#define SDVAE_ABL 0                // compile-time ablation mask of tuning builds: 2 no gather, 4 no MMAs, 16 no TMEM stores, 32 no g staging
#endif

__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}


// Two k-steps (K = 8 each) of one A chunk against the hi / lo images of g: six tcgen05.mma in one block, the B
// descriptors derived from the chunk's first one by immediate adds (1 KB = 64 descriptor units to the lo image,
// 2 KB = 128 to the next k-step).  The issuing thread is the kernel's critical path and competes for issue slots with
// five other warps of its SM sub-partition (profiles/r02_tile_kernel.md): every instruction saved here counts.
__device__ __forceinline__ void umma_bw_2k(uint32_t d_tmem, uint32_t a_hi, uint64_t bd_hi, uint32_t idesc, uint32_t acc_first) {
    asm volatile(
        "{\n.reg .pred p, q;\n.reg .b64 dl, dh1, dl1;\n.reg .b32 al, ah1, al1;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "setp.eq.b32 q, %4, %4;\n"
        "add.s64 dl, %2, 64;\n"
        "add.s64 dh1, %2, 128;\n"
        "add.s64 dl1, %2, 192;\n"
        "add.u32 al, %1, 32;\n"
        "add.u32 ah1, %1, 8;\n"
        "add.u32 al1, %1, 40;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], dl, %3, q;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], %2, %3, q;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], dh1, %3, q;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], dl1, %3, q;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al1], dh1, %3, q;\n}\n"
        ::"r"(d_tmem), "r"(a_hi), "l"(bd_hi), "r"(idesc), "r"(acc_first) : "memory");
}

__global__ void __launch_bounds__(kWThreads, 1)
bt_kernel(const TileBwArgs a) {
    const int S = a.S;
    const int NTS = a.nts;
    const int ROWS_BYTES = a.rcap * 128;
    const int STAGE_BYTES = ROWS_BYTES + S * 512;
    const int NBLK = (S * 32 + 1 + 127) >> 7;             // accumulator blocks (incl. the ones row)
    const int ONES_ROW = S * 32;                          // M row of the db accumulator
    const int K = S * 32;
    const int CPT = NBLK * 4;                             // A chunks per tile: (block, 32-row group)

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* G_s = smem;                                   // [2][16][hi 1 KB | lo 1 KB]  g tiles (MN-major, swizzled)
    uint8_t* T_s = G_s + 2 * umma::kGStage;                // [NTS] tile stages: rows | cell words
    float* Acc_s = reinterpret_cast<float*>(T_s + (size_t)NTS * STAGE_BYTES);   // [NBLK*32][128] drain accumulators (thread-private columns)
    uint64_t* bars = reinterpret_cast<uint64_t*>(Acc_s + (size_t)NBLK * 32 * 128);
    uint64_t* tile_full = bars;                            // [NTS] loader -> splitters
    uint64_t* tile_empty = bars + kTMaxStages;             // [NTS] splitters (one arrival per warp) -> loader
    uint64_t* a_full = bars + 2 * kTMaxStages;             // [5]   splitters (4 warps) -> MMA
    uint64_t* a_empty = a_full + kWAStages;                // [5]   MMA (commit) -> splitters
    uint64_t* g_full = a_empty + kWAStages;                // [2]   g warps (4) -> MMA
    uint64_t* g_empty = g_full + 2;                        // [2]   MMA (commit) -> g warps
    uint64_t* done_bar = g_empty + 2;                      // [2]   MMA (commit) -> drain, per accumulator set
    uint64_t* drained_bar = done_bar + 2;                  // [2]   drain (4 warps) -> MMA: the set may be restarted
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained_bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kTMaxStages; ++i) { mbar_init(tile_full + i, 1); mbar_init(tile_empty + i, 4 * kWSplitSets); }
        for (int i = 0; i < kWAStages; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(g_full + i, 4); mbar_init(g_empty + i, 1);
            mbar_init(done_bar + i, 1); mbar_init(drained_bar + i, 4);
        }
        fence_barrier_init();
    }
    if (warp == kWMmaWarp) {
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
    const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;
    const int T = a.flush;

    if (warp == kWMmaWarp) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t IDESC = umma::idesc_tf32_bmn(kBM, umma::kBwNT);
            const uint64_t g_desc = umma::smem_desc_mn_sw128(smem_u32(G_s));   // + (byte offset >> 4): the address field has room
            const uint32_t bar_full = smem_u32(a_full);
            int as = 0; uint32_t aph = 0;
            int tf = 0, nfl = 0;                           // tile index inside the flush group, flushes done
            bool ready = my_tiles > 0 && mbar_try_wait_a(bar_full, 0u);
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                const int gb = it & 1;
                const int ab = nfl & 1;                    // accumulator set of this flush group
                if (tf == 0 && nfl >= 2) {                 // previous use of this set drained?
                    mbar_wait(drained_bar + ab, (uint32_t)(((nfl >> 1) - 1) & 1));
                    tc_fence_after();
                }
                const bool last_of_group = (tf == T - 1) || (it == my_tiles - 1);
                mbar_wait(g_full + gb, (it >> 1) & 1);
#pragma unroll 1
                for (int c = 0; c < CPT; ++c) {
                    const int blk = c >> 2, r4 = c & 3;
                    if (!ready) {
                        int spins = 0;
                        while (!mbar_try_wait_a(bar_full + (uint32_t)as * 8u, aph)) { if (++spins > kSpinLimit) __trap(); }
                    }
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)((ab * NBLK + blk) * umma::kBwNT);
                    const uint32_t a_hi = tmem_base + (uint32_t)(kWAColBase + as * 64);   // lo columns: + 32
                    const uint64_t bd0 = g_desc + (uint64_t)((gb * umma::kGStage + r4 * 4 * 2048) >> 4);
                    uint64_t* const my_empty = a_empty + as;
                    if (++as == kWAStages) { as = 0; aph ^= 1; }
                    const bool more = c + 1 < CPT || it + 1 < my_tiles;
                    if (SDVAE_ABL & 4) ready = more && mbar_try_wait_a(bar_full + (uint32_t)as * 8u, aph);
                    else {
                        umma_bw_2k(d_tmem, a_hi, bd0, IDESC, (uint32_t)(tf | r4));
                        ready = more && mbar_try_wait_a(bar_full + (uint32_t)as * 8u, aph);
                        umma_bw_2k(d_tmem, a_hi + 16, bd0 + 256, IDESC, 1u);
                    }
                    umma_commit(my_empty);
                }
                umma_commit(g_empty + gb);
                if (last_of_group) { umma_commit(done_bar + ab); tf = 0; ++nfl; } else ++tf;
            }
        }
        __syncwarp();
    } else if (warp >= kWFirstLoadWarp) {
        // ================= loaders: warp lw owns tile stage lw (as gt_kernel's, forward plans only) =================
        const int lw = warp - kWFirstLoadWarp;
        if (lw < NTS) {
            const int q = lane & 7, rsub = lane >> 3;
            const uint32_t stage_a = smem_u32(T_s) + (uint32_t)lw * (uint32_t)STAGE_BYTES;
            // odd positions: high half first (the parity of staged row e = 32j + 4t + rsub is that of rsub)
            const uint32_t dst_rows = stage_a + (uint32_t)rsub * 128u + (((uint32_t)q * 16u) ^ ((uint32_t)(rsub & 1) << 6));
            const uint32_t dst_cell = stage_a + (uint32_t)ROWS_BYTES;
            const int n_cell16 = (S * 512) >> 4;
            constexpr int PV = kTMaxRcap / 32;
            long long t0 = (long long)blockIdx.x + (long long)lw * gridDim.x;
            int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
            uint32_t tph = 0;
#pragma unroll 1
            for (int it = lw; it < my_tiles; it += NTS) {
                PlanRegs<PV> now;
                plan_fetch(now, a.plan_cnt, a.plan_src, jt, 1, 0, a.rcap, rsub);
                const char* gb = reinterpret_cast<const char*>(a.in + (size_t)b * a.in_rows * a.in_ld + 4 * q);
                const char* cell_g = reinterpret_cast<const char*>(a.plan_cell + (size_t)jt * S * 128);
                const uint32_t row_bytes = (uint32_t)a.in_ld * 4u;
                mbar_wait_relaxed(tile_empty + lw, tph ^ 1);
#pragma unroll
                for (int j = 0; j < PV; ++j) {
                    if (32 * j < now.n) {
                        const uint32_t w[4] = {now.w[j].x, now.w[j].y, now.w[j].z, now.w[j].w};
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const uint32_t row = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                                         ::"r"(dst_rows + (uint32_t)(32 * j + 4 * t) * 128u), "l"(gb + (size_t)row * row_bytes));
                        }
                    }
                }
#pragma unroll 1
                for (int i = lane; i < n_cell16; i += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_cell + (uint32_t)i * 16u), "l"(cell_g + i * 16));
                cp_async_commit();
                cp_async_wait<0>();
                warp_arrive(tile_full + lw, lane);
                tph ^= 1;
                for (int k = 0; k < NTS; ++k) {
                    b += db; jt += djt;
                    if (jt >= a.L) { jt -= a.L; ++b; }
                }
            }
        }
    } else if (warp < kWFirstSplitWarp) {
        // ================= g staging (per tile), accumulator drain (per flush group) -- as bw_umma_kernel =================
        const int p = tid;                                        // 0..127
        int b = (int)blockIdx.x / a.L, jt = (int)blockIdx.x - b * a.L;
        const int n_real = a.n_real;
        const bool vec = (n_real == umma::kBwNT) && ((reinterpret_cast<uintptr_t>(a.g) & 15) == 0) && (a.g_ld % 4 == 0);
        const int q4 = warp & 3;
        float* P = a.part + (size_t)blockIdx.x * a.part_cta;
        float* Pb = a.part_b ? a.part_b + (size_t)blockIdx.x * a.partb_cta : nullptr;
        // The accumulators are drained into a per-CTA partial in SHARED memory (thread p owns column p of
        // Acc_s[NBLK*32][128]: no synchronisation, no bank conflicts, round-to-nearest adds in a fixed order) and
        // written to the CTA's global partial once at the end.  [Draining straight into L2 with RED.ADD cost
        // 96 x 128 reductions per drain, ~1.3 clk per lane and SM: 15.8 k clk per two tiles against 4.6 k clk of
        // MMAs -- it bounded both weight-gradient kernels, profiles/r02_dw.md.]
        for (int i = 0; i < NBLK * 32; ++i) Acc_s[i * 128 + p] = 0.f;
        auto drain = [&](int f) {
            const int ab = f & 1;
            mbar_wait(done_bar + ab, (uint32_t)((f >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int blk = 0; blk < NBLK; ++blk) {
                const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((ab * NBLK + blk) * umma::kBwNT);
                float v1[16], v2[16];
                tmem_ld16(t_row, v1);
                tmem_ld16(t_row + 16, v2);
                tmem_ld_wait();
                if (blk == NBLK - 1) {                            // last read of this set: hand it back
                    tc_fence_before();
                    warp_arrive(drained_bar + ab, lane);
                }
                float* acc = Acc_s + (size_t)blk * 32 * 128 + p;
                if (!(SDVAE_ABL & 64))
#pragma unroll
                for (int j = 0; j < 16; ++j) { acc[j * 128] += v1[j]; acc[(16 + j) * 128] += v2[j]; }
            }
        };
        auto write_out = [&]() {
#pragma unroll 1
            for (int blk = 0; blk < NBLK; ++blk) {
                const int row = blk * 128 + q4 * 32 + lane;       // M row = s*32 + c
                // M row = s*32 + c -> column s*in_ld + c of the layer's [n, S*C_in] weight gradient
                float* dst = row < K ? P + (row >> 5) * a.in_ld + (row & 31) : (row == ONES_ROW ? Pb : nullptr);
                const size_t ld = row < K ? (size_t)a.part_ld : (size_t)1;
                const float* acc = Acc_s + (size_t)blk * 32 * 128 + p;
                if (dst)
                    for (int j = 0; j < n_real; ++j) dst[(size_t)j * ld] = acc[j * 128];
            }
        };
        int tf = 0, nfl = 0;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int gb = it & 1;
            uint8_t* gs = G_s + gb * umma::kGStage;
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
            if (!(SDVAE_ABL & 32))
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = vec ? (p >> 3) + 16 * i : p;
                const int q = vec ? (p & 7) : i;
                float4 hi, lo;
                split_tf32f(v[i].x, hi.x, lo.x); split_tf32f(v[i].y, hi.y, lo.y);
                split_tf32f(v[i].z, hi.z, lo.z); split_tf32f(v[i].w, hi.w, lo.w);
                uint8_t* dst = gs + umma::g_off(m, q);
                *reinterpret_cast<float4*>(dst) = hi;
                *reinterpret_cast<float4*>(dst + 1024) = lo;
            }
            fence_async_smem();
            warp_arrive(g_full + gb, lane);
            b += db; jt += djt;
            if (jt >= a.L) { jt -= a.L; ++b; }
            // tile `it` is staged; if tile it-1 closed a flush group, drain it now (the MMA warp is waiting)
            if (tf == 0 && it > 0) { drain(nfl); ++nfl; }
            tf = (tf == T - 1) ? 0 : tf + 1;
        }
        if (my_tiles > 0) drain(nfl);                             // the last group always ends with a flush
        write_out();
    } else {
        // ================= splitters (transposing): thread = M row (slot s, channel c), K = 32 tile rows =================
        const int set = (warp - kWFirstSplitWarp) >> 2;
        const int q4 = warp & 3;
        const uint32_t T_a = smem_u32(T_s);
        const uint32_t bar_tile_full = smem_u32(tile_full), bar_tile_empty = smem_u32(tile_empty);
        const uint32_t bar_a_full = smem_u32(a_full), bar_a_empty = smem_u32(a_empty);
        // channel `lane` of a staged row whose LOW half starts at byte offset off: (off ^ cx) + cy
        const uint32_t cx = lane >= 16 ? 64u : 0u, cy = (uint32_t)(lane & 15) * 4u;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)kWAColBase;
        int jt = (int)blockIdx.x % a.L;
        int ts = 0; uint32_t tph = 0;
        uint32_t stage_a = T_a;
        int as = set; uint32_t aph = 0;                            // A stage / phase of chunk g = it*CPT + c (g % 5, (g / 5) & 1)
        int first = set;                                           // first chunk of this set in the current tile
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int nvalid = min(kBM, a.out_rows - jt * kBM);
            mbar_wait_a<64>(bar_tile_full + (uint32_t)ts * 8u, tph);
#pragma unroll 1
            for (int c = first; c < CPT; c += kWSplitSets) {
                const int blk = c >> 2, r4 = c & 3;
                const int s = blk * 4 + q4;
                const int row0 = blk * 128 + q4 * 32;              // first M row of this warp
                float v[32];
                bool have = false;
                if ((SDVAE_ABL & 2) && s < S) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (float)(lane + j);
                    have = true;
                } else if (s < S) {
                    // the 32 cell words of (slot s, tile rows 32*r4 ..): word of row 32*r4 + j at (j & 7)*4 + (j >> 3)
                    const uint32_t wbase = stage_a + (uint32_t)ROWS_BYTES + (uint32_t)(s * 128 + r4 * 32) * 4u;
                    const uint32_t rbase = stage_a + cy;
                    const bool whole = 32 * r4 + 32 <= nvalid;    // warp-uniform: all 32 rows of the group exist
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint4 w4 = lds128u(wbase + (uint32_t)i * 16u);          // warp-uniform address: broadcast
                        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) v[i + 8 * h] = lds32(rbase + ((w[h] & 0xffffu) ^ cx));
                    }
                    if (!whole) {                                  // rows past the mesh: plan word = staged row 0 (valid), value dropped
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 32 * r4 + j < nvalid ? v[j] : 0.f;
                    }
                    have = true;
                } else if (row0 <= ONES_ROW && ONES_ROW < row0 + 32) {
                    // the warp that owns the ones row: A^T[ONES_ROW, m] = 1 for valid rows, everything else 0
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (row0 + lane == ONES_ROW && 32 * r4 + j < nvalid) ? 1.f : 0.f;
                    have = true;
                }
                mbar_wait_a<128>(bar_a_empty + (uint32_t)as * 8u, aph ^ 1);
                __syncwarp();
                if (have) {                                        // (warps past the ones row leave their TMEM lanes alone)
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float2 v2 = make_float2(v[j], v[j + 1]);
                        const float2 hi = make_float2(__uint_as_float(__float_as_uint(v2.x) & 0xffffe000u),
                                                      __uint_as_float(__float_as_uint(v2.y) & 0xffffe000u));
                        const float2 l2 = sub2(v2, hi);
                        v[j] = hi.x; v[j + 1] = hi.y; lo[j] = l2.x; lo[j + 1] = l2.y;
                    }
                    tc_fence_after();
                    const uint32_t t_a = t_lane + (uint32_t)(as * 64);
                    if (!(SDVAE_ABL & 16)) {
                        tmem_st32(t_a, v);
                        tmem_st32(t_a + 32, lo);
                        tmem_st_wait();
                    } else if (lo[3] + v[5] == 12345.678f) a.part[0] = 1.f;
                }
                tc_fence_before();
                if (lane == 0) mbar_arrive_a(bar_a_full + (uint32_t)as * 8u);
                as += kWSplitSets;                                  // next chunk of this set: g + 4
                if (as >= kWAStages) { as -= kWAStages; aph ^= 1; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bar_tile_empty + (uint32_t)ts * 8u);   // this warp is done with the tile stage
            stage_a += (uint32_t)STAGE_BYTES;
            if (++ts == NTS) { ts = 0; tph ^= 1; stage_a = T_a; }
            first -= CPT % kWSplitSets;                            // CPT is a multiple of 4: `first` stays = set
            if (first < 0) first += kWSplitSets;
            jt += djt; if (jt >= a.L) jt -= a.L;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace tile
}  // namespace sdvae
