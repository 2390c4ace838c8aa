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
//   * the splitters gather from shared memory WITHOUT bank conflicts for arbitrary rows: they write the TMEM
//     A operand with tcgen05.st.16x256b (four threads per row), the four threads of a row read one 64-byte half
//     of it with one LDS.128, and the two rows of an 8-lane phase read opposite halves.  The price is a fixed
//     permutation of the 32 channels of a K chunk (kperm below), applied to the weight image by the packer.
// Precision, weight image, MMA form (TS: A from TMEM, B = resident weight image), epilogue: as gc_umma_kernel
// (error-compensated 3xTF32, fp32 accumulation in TMEM).
//
// Tile plan (per tile t of 128 output rows; all tables static, built once on the host):
//   plan_cnt [L]             distinct source rows of the tile (<= rcap)
//   plan_src [L, rcap/2]     those rows, 16-bit pairs in loader-lane order (umma::plan_fetch, S = 1)
//   plan_cell[L, S*128]      one word per (slot s, tile row r), stored at  s*128 + (r>>5)*32 + (r&7)*4 + ((r>>3)&3)
//                            (the four words of a splitter thread are one 16-byte load):
//                              bits 0..8 position of the cell's first row in the tile's list, bits 9..13 rows in the
//                              cell, bits 14.. offset of its 2nd, 3rd, ... rows in plan_ext
//   plan_ext [L, ecap]       16-bit positions of the 2nd, 3rd ... rows of the cells (RAGGED plans only)
#pragma once
#include "spiral_conv_umma.cuh"

namespace sdvae {
namespace tile {

using namespace umma;

constexpr int kTEpilogueWarps = 4;
constexpr int kTSplitSets = 4;
constexpr int kTSplitWarps = 4 * kTSplitSets;
constexpr int kTMaxStages = 3;                                  // tile-stage ring depth limit (one loader warp each)
constexpr int kTFirstSplitWarp = kTEpilogueWarps;               // 4
constexpr int kTFirstLoadWarp = kTFirstSplitWarp + kTSplitWarps;  // 20
constexpr int kTMmaWarp = kTFirstLoadWarp + kTMaxStages;        // 23
constexpr int kTThreads = (kTMmaWarp + 1) * 32;                 // 768 -> 80 registers per thread at launch
// setmaxnreg budgets per warpgroup: 128*64 + 512*88 + 128*64 = 61440 = 768*80
constexpr int kTRegsEpilogue = 64, kTRegsSplit = 88, kTRegsLoad = 64;
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

template <bool RAGGED>
__global__ void __launch_bounds__(kTThreads, 1)
gt_kernel(const TileArgs a) {
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
        for (int i = 0; i < kTMaxStages; ++i) { mbar_init(tile_full + i, 1); mbar_init(tile_empty + i, 4 * NCH); }
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
        reg_dec<kTRegsLoad>();
        if (warp == kTMmaWarp) {
            // ================= MMA issuer: one wait + 24 MMAs + one commit per stage of three chunks =================
            if (elect_one()) {
                constexpr uint32_t IDESC1 = idesc_tf32(kBM, 2 * kTNT);
                constexpr uint32_t IDESC2 = idesc_tf32(kBM, kTNT);
                const uint64_t desc0 = smem_desc_sw128(smem_u32(B_s));
                const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
                int st = 0; uint32_t sph = 0;
#pragma unroll 1
                for (int it = 0; it < my_tiles; ++it) {
                    const int acc = it & 1;
                    mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * kTNT);
#pragma unroll 1
                    for (int c0 = 0; c0 < NCH; c0 += kTChunksPerStage) {
                        mbar_wait(a_full + st, sph);
                        tc_fence_after();
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
            }
            __syncwarp();
        } else {
            // ================= loaders: warp lw owns tile stage lw, one whole tile per pass =================
            const int lw = warp - kTFirstLoadWarp;
            if (lw < NTS) {
                const int q = lane & 7, rsub = lane >> 3;
                uint8_t* stage = T_s + (size_t)lw * STAGE_BYTES;
                const uint32_t dst_rows = smem_u32(stage) + (uint32_t)rsub * 128u + (uint32_t)q * 16u;
                const uint32_t dst_cell = smem_u32(stage) + (uint32_t)ROWS_BYTES;
                const int n_cell16 = (CELL_BYTES + a.ecap * 2) >> 4;       // cell words and ext are contiguous in the stage
                constexpr int PV = kTMaxRcap / 32;
                long long t0 = (long long)blockIdx.x + (long long)lw * gridDim.x;
                int b = (int)(t0 / a.L), jt = (int)(t0 - (long long)b * a.L);
                uint32_t tph = 0;
#pragma unroll 1
                for (int it = lw; it < my_tiles; it += NTS) {
                    PlanRegs<PV> now;
                    plan_fetch(now, a.plan_cnt, a.plan_src, jt, 1, 0, a.rcap, rsub);
                    const float* base = a.in + (size_t)b * a.in_rows * 32 + 4 * q;
                    const char* cell_g = reinterpret_cast<const char*>(a.plan_cell + (size_t)jt * S * 128);
                    const char* ext_g = RAGGED ? reinterpret_cast<const char*>(a.plan_ext + (size_t)jt * a.ecap) : nullptr;
                    mbar_wait_relaxed(tile_empty + lw, tph ^ 1);
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
                    cp_async_wait<0>();
                    warp_arrive(tile_full + lw, lane);
                    tph ^= 1;
                    for (int k = 0; k < NTS; ++k) {
                        b += db; jt += djt;
                        if (jt >= a.L) { jt -= a.L; ++b; }
                    }
                }
            }
        }
    } else if (warp < kTFirstSplitWarp) {
        reg_dec<kTRegsEpilogue>();
        // ================= epilogue (as gc_umma_kernel's forward epilogue, rows staged for coalesced stores) ======
        const int q4 = warp & 3;
        const int EPI = a.epi;
        const int ldo = a.ldo;
        const bool has_bias = (EPI == EPI_BIAS || EPI == EPI_BIAS_ELU) && a.bias != nullptr;
        int b = (int)blockIdx.x / a.L, jt = (int)blockIdx.x - b * a.L;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            mbar_wait_relaxed(t_full + acc, (it >> 1) & 1);
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
                for (int j = 0; j < 16; ++j) v[j] += d2[j];
                if (has_bias) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bv = ldg4(a.bias + c0 + j);
                        v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                    }
                }
                if (EPI == EPI_BIAS_ELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = elu_fast(v[j]);
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
    } else {
        reg_inc<kTRegsSplit>();
        // ================= splitters =================
        // set k takes the chunks g = k (mod 4) of the CTA's chunk sequence (tile iteration, slot); its four warps
        // own the four TMEM lane quarters.  Thread (l4 = lane >> 2, qq = lane & 3) serves the tile rows
        // 32*q4 + 16*g + 8*h + l4 (g, h in {0, 1}): for each it reads the two 16-byte pieces qq and qq + 4 of the
        // staged row(s) of the cell -- rows with even l4 the low 64-byte half first, rows with odd l4 the high half
        // first, so the eight lanes of an LDS.128 phase (two rows) always cover all 32 banks.
        const int set = (warp - kTFirstSplitWarp) >> 2;
        const int q4 = warp & 3;
        const int l4 = lane >> 2, qq = lane & 3;
        const bool odd = (l4 & 1) != 0;
        const uint32_t offX = (uint32_t)qq * 16u + (odd ? 64u : 0u), offY = (uint32_t)qq * 16u + (odd ? 0u : 64u);
        const uint32_t cell_off = (uint32_t)ROWS_BYTES + (uint32_t)(q4 * 32 + l4 * 4) * 4u;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)kTAccCols;
        int g = set, ch = set, it = 0;
        while (ch >= NCH) { ch -= NCH; ++it; }
#pragma unroll 1
        while (g < G) {
            const int ts = it % NTS;
            const uint32_t tph = (uint32_t)((it / NTS) & 1);
            const int sg = g / kTChunksPerStage;                 // stage round of this chunk
            const int st = sg & 1;
            const uint32_t sph = (uint32_t)((sg >> 1) & 1);
            const int sub = g - sg * kTChunksPerStage;
            mbar_wait_relaxed(a_empty + st, sph ^ 1);            // order: the TMEM stage first, then the tile stage
            mbar_wait_relaxed(tile_full + ts, tph);
            const uint8_t* stage = T_s + (size_t)ts * STAGE_BYTES;
            const uint4 cw = *reinterpret_cast<const uint4*>(stage + cell_off + (uint32_t)ch * 512u);
            const uint32_t words[4] = {cw.x, cw.y, cw.z, cw.w};
            const uint32_t t_a = t_lane + (uint32_t)(st * kTStageCols + sub * 64);
            tc_fence_after();
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
                float r[32];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t w = words[gg * 2 + h];
                    const uint8_t* row = stage + (w & 0x1ffu) * 128u;
                    float4 X = *reinterpret_cast<const float4*>(row + offX);
                    float4 Y = *reinterpret_cast<const float4*>(row + offY);
                    if (RAGGED) {
                        const int cnt = (int)((w >> 9) & 0x1fu);
                        if (cnt == 0) { X = make_float4(0.f, 0.f, 0.f, 0.f); Y = X; }
                        const uint16_t* ext = reinterpret_cast<const uint16_t*>(stage + ROWS_BYTES + CELL_BYTES) + (w >> 14);
#pragma unroll 1
                        for (int e = 1; e < cnt; ++e) {           // in-order sum: deterministic scatter-add
                            const uint8_t* row2 = stage + (uint32_t)ext[e - 1] * 128u;
                            const float4 X2 = *reinterpret_cast<const float4*>(row2 + offX);
                            const float4 Y2 = *reinterpret_cast<const float4*>(row2 + offY);
                            X.x += X2.x; X.y += X2.y; X.z += X2.z; X.w += X2.w;
                            Y.x += Y2.x; Y.y += Y2.y; Y.z += Y2.z; Y.w += Y2.w;
                        }
                    }
                    const float4 P = odd ? Y : X;                // piece qq     : channels 4qq .. 4qq+3   -> n = 0, 1
                    const float4 Q = odd ? X : Y;                // piece qq + 4 : channels 16+4qq ..      -> n = 2, 3
                    const float v8[8] = {P.x, P.y, P.z, P.w, Q.x, Q.y, Q.z, Q.w};
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
#pragma unroll
                        for (int bb = 0; bb < 2; ++bb) {
                            float hi, lo;
                            split_tf32f(v8[2 * n + bb], hi, lo);
                            r[4 * n + 2 * h + bb] = hi;
                            r[16 + 4 * n + 2 * h + bb] = lo;
                        }
                    }
                }
                tmem_st_16x256b_x8(t_a + ((uint32_t)(16 * gg) << 16), r);
            }
            warp_arrive(tile_empty + ts, lane);      // this warp's reads of the tile stage for this chunk are done
            tmem_st_wait();
            tc_fence_before();
            warp_arrive(a_full + st, lane);
            g += kTSplitSets; ch += kTSplitSets;
            while (ch >= NCH) { ch -= NCH; ++it; }
        }
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
