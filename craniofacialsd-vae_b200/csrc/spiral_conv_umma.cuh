// Spiral convolution on the 5th-generation tensor cores (tcgen05 + TMEM), for the wide
// layers (C_in in {32, 64}: K = S*C_in = 288 / 576).  Reference op: model.py:27-41 + F.elu
// (model.py:68,84); its backward-to-input is the same contraction over the inverse table.
//
//   y[m, n] = epi( sum_{s,c} A[m, s*KS + c] * W[n, s*KS + c] ),   A[m, s*KS+c] = x[src(m,s), c]
//
// Precision: error-compensated 3xTF32.  Every fp32 operand v is split into
//   hi = v with the low 13 mantissa bits cleared (exact in TF32),  lo = v - hi (exact in fp32)
// and the product is accumulated in fp32 (TMEM) as  A_hi*W_hi + A_hi*W_lo + A_lo*W_hi.
// The first two terms are ONE tcgen05.mma with N = 2*NT (the weight image stacks the W_hi
// rows over the W_lo rows; the two halves land in adjacent TMEM column ranges and are added
// in the epilogue), the third is a second MMA with N = NT onto the first half.
//
// Persistent, warp-specialised CTA (one per SM, 416 threads):
//   warp 0      : TMEM allocation, MMA issue (one elected lane), tcgen05.commit -> mbarriers
//   warps 1..4  : epilogue: tcgen05.ld accumulator -> registers -> bias/ELU/ELU'-gate -> global
//   warps 5..12 : producers: gather rows of x (LDG.128, 8 lanes per 128-byte row), split hi/lo in
//                 registers, store both into the 128B-swizzled K-major UMMA tile in shared memory
// Pipelines: A-tile ring (full/empty mbarriers), double-buffered TMEM accumulator
// (tmem_full/tmem_empty), static round-robin tile schedule.  The weight image (already split,
// permuted into the swizzled layout by umma_pack_weights_kernel) stays resident in shared
// memory for the whole kernel.
//
// Shared-memory operand layout (both A and B): K-major, SWIZZLE_128B.  One 32-float K chunk
// of a row is 128 bytes = one swizzle row; 8 rows form a 1024-byte atom (SBO = 1024); the
// 16-byte column j of row r sits at  r*128 + ((j ^ (r & 7)) * 16).  One tcgen05.mma consumes
// K = 8 tf32 (32 bytes): the descriptor start address advances by 32 bytes per k-step.
#pragma once
#include "common.cuh"
#include "spiral_conv.cuh"

namespace sdvae {
namespace umma {

constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueWarps = 4;
constexpr int kFirstEpilogueWarp = 1;
constexpr int kFirstProducerWarp = kFirstEpilogueWarp + kEpilogueWarps;     // 5
constexpr int kThreads = (kFirstProducerWarp + kProducerWarps) * 32;        // 416
constexpr int kBM = 128;                                                    // rows per tile (UMMA M)
constexpr int kPrefetch = 2;                                                // producer register prefetch distance (chunks)
constexpr long long kSpinLimit = 4000000000LL;                              // cycles before a stuck wait traps

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread for the CTA.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp <- lane (base_lane + t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor: D = F32, A = B = TF32, both K-major, dense.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_tf32f(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}
// byte offset of the 16-byte column q of row r inside a [rows x 128 B] SWIZZLE_128B tile
__host__ __device__ __forceinline__ int sw128_off(int r, int q) { return r * 128 + ((q ^ (r & 7)) << 4); }

// ---- weight image ----------------------------------------------------------------------------
// img[chunk][j][32]:  j < NT -> hi part of W row j,  j >= NT -> lo part of row j-NT; rows >= n_real
// are zero.  `transposed` selects the backward-to-input weight  Wt[c, s*Cout + o] = W[o, s*Cin + c]
// read straight from the forward weight (so no separate transpose pass is needed):
//   forward   : n = output channel, k = s*KS + c      -> W[n*ldw + k]                 (KS = Cin)
//   transposed: n = input channel c, k = s*KS + o     -> W[o*ldw + s*n_real_src + n]  (KS = Cout)
struct PackArgs {
    const float* W;
    float* img;
    int NT, KS, S, n_real, ldw, transposed, cin;   // cin: C_in of the forward layer (transposed only)
};

__global__ void umma_pack_weights_kernel(const PackArgs a) {
    const int K = a.S * a.KS;
    const int total = (K / 32) * 2 * a.NT * 32;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int kk = t & 31;
        const int j = (t >> 5) % (2 * a.NT);
        const int ch = (t >> 5) / (2 * a.NT);
        const int n = j % a.NT, part = j / a.NT;
        const int k = ch * 32 + kk;
        float w = 0.f;
        if (n < a.n_real) {
            if (!a.transposed) {
                w = a.W[(size_t)n * a.ldw + k];
            } else {
                const int s = k / a.KS, o = k - s * a.KS;
                w = a.W[(size_t)o * a.ldw + s * a.cin + n];
            }
        }
        float hi, lo;
        split_tf32f(w, hi, lo);
        const int off = ch * (2 * a.NT * 128) + sw128_off(j, kk >> 2) + (kk & 3) * 4;
        a.img[off >> 2] = part ? lo : hi;
    }
}

// ---- main kernel -----------------------------------------------------------------------------
template <int KS, int NT>
struct UmmaCfg {
    static constexpr int CPS = KS / 32;                  // 32-wide chunks per spiral slot
    static constexpr int A_STAGE = 2 * kBM * 128;        // hi tile + lo tile
    static constexpr int B_CHUNK = 2 * NT * 128;
    static constexpr int TMEM_COLS = 4 * NT < 32 ? 32 : 4 * NT;   // two accumulators of 2*NT columns
    static int stages(int S) {
        const int budget = 225 * 1024 - 2048 - S * CPS * B_CHUNK - 2 * kBM * (S + 2) * 4;
        int st = budget / A_STAGE;
        return st > 6 ? 6 : st;
    }
    static size_t smem_bytes(int S, int nst) {
        return 1024 /*align slack*/ + (size_t)S * CPS * B_CHUNK + (size_t)nst * A_STAGE +
               (size_t)2 * kBM * (S + 2) * 4 + 1024 /*barriers*/;
    }
};

struct UmmaArgs {
    GcArgs g;            // in / idx / cell_ptr / cell_src / bias / gate / out / M / in_rows / Vout / S / ldo / n_real / epi
    const float* wimg;   // packed weight image (umma_pack_weights_kernel)
    int nstages;
    int ntiles;
};

template <int KS, int NT, bool RAGGED>
__global__ void __launch_bounds__(kThreads, 1)
gc_umma_kernel(const UmmaArgs ua) {
    using Cfg = UmmaCfg<KS, NT>;
    constexpr int CPS = Cfg::CPS, A_STAGE = Cfg::A_STAGE, B_CHUNK = Cfg::B_CHUNK;
    const GcArgs& a = ua.g;
    const int S = a.S;
    const int NCH = S * CPS;
    const int NST = ua.nstages;
    const int TS = S + 1;                                  // table row stride (ragged keeps S+1 cell bounds)

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* B_s = smem;                                   // [NCH][2NT][128 B]
    uint8_t* A_s = B_s + (size_t)NCH * B_CHUNK;            // [NST][hi 16 KB | lo 16 KB]
    int* T_s = reinterpret_cast<int*>(A_s + (size_t)NST * A_STAGE);      // [2][kBM][TS] source-row tables
    int* R_s = T_s + 2 * kBM * TS;                         // [2][kBM] ragged: first row of the mesh in `in`
    uint64_t* bars = reinterpret_cast<uint64_t*>(R_s + 2 * kBM);
    bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 15) & ~(uintptr_t)15);
    uint64_t* full_bar = bars;                             // [NST]  producers -> MMA
    uint64_t* empty_bar = bars + 8;                        // [NST]  MMA (commit) -> producers
    uint64_t* tfull_bar = bars + 16;                       // [2]    MMA (commit) -> epilogue
    uint64_t* tempty_bar = bars + 18;                      // [2]    epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(full_bar + i, kProducerThreads); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, kEpilogueWarps * 32); }
        fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    {   // resident weight image
        const int n16 = NCH * B_CHUNK / 16;
        const float4* src = reinterpret_cast<const float4*>(ua.wimg);
        float4* dst = reinterpret_cast<float4*>(B_s);
        for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = ua.ntiles;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t IDESC1 = idesc_tf32(kBM, 2 * NT);
            constexpr uint32_t IDESC2 = idesc_tf32(kBM, NT);
            const uint32_t a_base = smem_u32(A_s), b_base = smem_u32(B_s);
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = it & 1;
                mbar_wait(tempty_bar + acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * NT);
                for (int ch = 0; ch < NCH; ++ch) {
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t a_hi = a_base + stage * A_STAGE, a_lo = a_hi + kBM * 128;
                    const uint32_t b_ch = b_base + ch * B_CHUNK;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = smem_desc_sw128(b_ch + k * 32);
                        umma_tf32(d_tmem, smem_desc_sw128(a_hi + k * 32), bd, IDESC1, (ch | k) != 0);
                        umma_tf32(d_tmem, smem_desc_sw128(a_lo + k * 32), bd, IDESC2, 1u);
                    }
                    umma_commit(empty_bar + stage);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar + acc);
            }
        }
        __syncwarp();
    } else if (warp < kFirstProducerWarp) {
        // ================= epilogue =================
        const int q4 = warp & 3;                                  // TMEM lane quarter this warp may read
        const int EPI = a.epi;
        const int n_real = a.n_real, ldo = a.ldo;
        const bool vec_ok = (n_real == NT) && ((ldo & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) &&
                            (EPI != EPI_GATE || (reinterpret_cast<uintptr_t>(a.gate) & 15) == 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int acc = it & 1;
            mbar_wait(tfull_bar + acc, (it >> 1) & 1);
            tc_fence_after();
            const long long m = (long long)tile * kBM + q4 * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * 2 * NT);
#pragma unroll
            for (int c0 = 0; c0 < NT; c0 += 16) {
                float d1[16], d2[16];
                tmem_ld16(t_row + c0, d1);
                tmem_ld16(t_row + NT + c0, d2);
                tmem_ld_wait();
                if (c0 + 16 >= NT) {            // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(tempty_bar + acc);
                }
                if (m < a.M) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = d1[j] + d2[j];
                        const int col = c0 + j;
                        if ((EPI == EPI_BIAS || EPI == EPI_BIAS_ELU) && a.bias && col < n_real) v[j] += __ldg(a.bias + col);
                        if (EPI == EPI_BIAS_ELU) v[j] = elu_f(v[j]);
                    }
                    const size_t off = (size_t)m * ldo + c0;
                    if (vec_ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            if (EPI == EPI_GATE) {
                                const float4 gt = ldg4(a.gate + off + j);
                                o.x *= elu_grad_from_out(gt.x); o.y *= elu_grad_from_out(gt.y);
                                o.z *= elu_grad_from_out(gt.z); o.w *= elu_grad_from_out(gt.w);
                            }
                            *reinterpret_cast<float4*>(a.out + off + j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (c0 + j < n_real) {
                                float o = v[j];
                                if (EPI == EPI_GATE) o *= elu_grad_from_out(__ldg(a.gate + off + j));
                                a.out[off + j] = o;
                            }
                        }
                    }
                }
            }
        }
    } else {
        // ================= producers =================
        const int p = tid - kFirstProducerWarp * 32;              // 0..255
        const int q = p & 7, r0 = p >> 3;                         // 16-byte column, first row (rows r0 + 32*i)

        // source-row table of one tile:  uniform: T[lr][s] = absolute row of x (or -1)
        //                                ragged : T[lr][s] = first entry of cell (r,s), T[lr][S] = end; R[lr] = mesh base row
        auto table_load = [&](int tile, int (&reg)[6]) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int e = p + i * kProducerThreads;
                int val = -1;
                if (e < kBM * TS) {
                    const int lr = e / TS, s = e - lr * TS;
                    const long long m = (long long)tile * kBM + lr;
                    if (m < a.M) {
                        const int b = (int)(m / a.Vout);
                        const int r = (int)(m - (long long)b * a.Vout);
                        if (!RAGGED) {
                            if (s < S) val = b * a.in_rows + __ldg(a.idx + r * S + s);
                        } else {
                            val = __ldg(a.cell_ptr + r * S + s);
                        }
                    } else if (RAGGED) {
                        val = 0;
                    }
                }
                reg[i] = val;
            }
        };
        auto table_store = [&](int buf, int tile, const int (&reg)[6]) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int e = p + i * kProducerThreads;
                if (e < kBM * TS) T_s[buf * kBM * TS + e] = reg[i];
            }
            if (RAGGED && p < kBM) {
                const long long m = (long long)tile * kBM + p;
                R_s[buf * kBM + p] = (m < a.M) ? (int)(m / a.Vout) * a.in_rows : 0;
            }
        };
        auto gather = [&](int tb, int ch, float4 (&v)[4]) {
            const int s = ch / CPS, h = ch - s * CPS;
            const int* T = T_s + tb * kBM * TS;
            if (!RAGGED) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int src = T[(r0 + 32 * i) * TS + s];
                    v[i] = src >= 0 ? ldg4(a.in + (size_t)src * KS + h * 32 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                int e0[4], e1[4], s0[4];
                const float* rb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int lr = r0 + 32 * i;
                    e0[i] = T[lr * TS + s];
                    e1[i] = T[lr * TS + s + 1];
                    rb[i] = a.in + (size_t)R_s[tb * kBM + lr] * KS + h * 32 + 4 * q;
                    s0[i] = e1[i] > e0[i] ? __ldg(a.cell_src + e0[i]) : -1;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[i] = s0[i] >= 0 ? ldg4(rb[i] + (size_t)s0[i] * KS) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 4; ++i)          // cells fed by several output rows: in-order sum (deterministic)
                    for (int e = e0[i] + 1; e < e1[i]; ++e) {
                        const float4 t = ldg4(rb[i] + (size_t)__ldg(a.cell_src + e) * KS);
                        v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
                    }
            }
        };

        // table of the first tile
        {
            int reg[6];
            table_load((int)blockIdx.x, reg);
            table_store(0, (int)blockIdx.x, reg);
            named_bar_sync(1, kProducerThreads);
        }

        // register ring of kPrefetch + 1 = 3 chunks; NCH % 3 == 0 (checked by the host) keeps the
        // slot of chunk ch equal to ch % 3 in every tile, so all indices are compile-time
        static_assert(kPrefetch == 2, "producer ring is written for a prefetch distance of 2");
        float4 pre[3][4];
        gather(0, 0, pre[0]);
        gather(0, 1, pre[1]);

        int stage = 0; uint32_t phase = 0;
        int tb = 0;                              // table buffer of the tile whose chunks are being stored
        int nreg[6];
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const bool has_next = it + 1 < my_tiles;
            for (int ch0 = 0; ch0 < NCH; ch0 += 3) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int ch = ch0 + u;
                    // table of the next tile: loads at chunk 0, stores at chunk 1 between two producer
                    // barriers (the first: nobody still reads the buffer being overwritten; the second: published)
                    if (ch == 0 && has_next) table_load(tile + (int)gridDim.x, nreg);
                    if (ch == 1) {
                        named_bar_sync(1, kProducerThreads);
                        if (has_next) table_store(tb ^ 1, tile + (int)gridDim.x, nreg);
                        named_bar_sync(1, kProducerThreads);
                    }
                    // prefetch chunk ch + 2 (possibly of the next tile) into the free register slot
                    {
                        int pch = ch + 2, ptb = tb;
                        bool ok = true;
                        if (pch >= NCH) { pch -= NCH; ptb ^= 1; ok = has_next; }
                        if (ok) gather(ptb, pch, pre[(u + 2) % 3]);
                    }
                    // split + store chunk ch
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    {
                        uint8_t* hi_t = A_s + (size_t)stage * A_STAGE;
                        uint8_t* lo_t = hi_t + kBM * 128;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 x4 = pre[u][i];
                            float4 hi, lo;
                            split_tf32f(x4.x, hi.x, lo.x); split_tf32f(x4.y, hi.y, lo.y);
                            split_tf32f(x4.z, hi.z, lo.z); split_tf32f(x4.w, hi.w, lo.w);
                            const int off = sw128_off(r0 + 32 * i, q);
                            *reinterpret_cast<float4*>(hi_t + off) = hi;
                            *reinterpret_cast<float4*>(lo_t + off) = lo;
                        }
                    }
                    fence_async_smem();
                    mbar_arrive(full_bar + stage);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
            tb ^= 1;
        }
    }

    // ---- teardown ----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace umma
}  // namespace sdvae
