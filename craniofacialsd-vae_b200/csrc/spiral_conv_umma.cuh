// Spiral convolution on the 5th-generation tensor cores (tcgen05 + TMEM), for the wide
// layers (channels per slot KS in {32, 64}: K = S*KS = 288 / 576).  Reference op:
// model.py:27-41 + F.elu (model.py:68,84); its backward-to-input (autograd of model.py:34,40)
// is the same contraction over the inverse table.
//
//   y[m, n] = epi( sum_{s,c} A[m, s*KS + c] * W[n, s*KS + c] ),
//   A[m, s*KS + c] = sum_{r in cell(m, s)} x[r, c]        (one source row per cell in the forward pass)
//
// Precision: error-compensated 3xTF32.  Every fp32 operand v is split into
//   hi = v with the low 13 mantissa bits cleared (exact in TF32),  lo = v - hi (exact in fp32)
// and the product is accumulated in fp32 (TMEM) as  A_hi*W_hi + A_hi*W_lo + A_lo*W_hi.
// The first two terms are ONE tcgen05.mma with N = 2*NT (the weight image stacks the W_hi rows
// over the W_lo rows; the halves land in adjacent TMEM column ranges and are added in the
// epilogue), the third is a second MMA with N = NT onto the first half.
//
// The gather is driven by a host-built TILE PLAN (tc_plan.h): for every tile of 128 output
// rows of a mesh and every spiral slot, the list of source rows to stage and, per output row,
// the [start, count) range of staged rows whose sum is that row's A-operand cell.
//
// Persistent, warp-specialised CTA (one per SM, 28 warps; register budgets moved between the
// roles with setmaxnreg):
//   warps 0..3   : epilogue: tcgen05.ld accumulator -> bias/ELU/ELU'-gate -> global
//   warps 4..19  : splitters, four sets of four warps (warp % 4 = TMEM lane quarter), set k takes the chunks
//                  g = k (mod 4): staged rows (LDS.128, conflict-free through the 128B swizzle) -> in-order
//                  sum (backward plans) -> hi/lo split in registers -> tcgen05.st into the TMEM A ring.
//                  ncu (profiles/r01_gc_umma_stalls.txt) showed this role latency-bound with two sets:
//                  one instruction per ~13 clk per warp, nothing else above 55 % -- hence four sets and a
//                  straight-line forward variant (one staged row per tile row, no cell table)
//   warps 20..26 : loaders: cp.async 16 B (8 lanes per 128-byte row piece) into the raw shared-memory ring,
//                  one chunk in flight per warp, no register staging
//   warp 27      : TMEM allocation, MMA issue (one elected lane), tcgen05.commit -> mbarriers
// A never passes through shared memory on its way into the MMA (TS form: A from TMEM, B from
// shared memory), so no generic->async proxy fence sits on the gather path.  The weight image
// (split and laid out by umma_pack_weights_kernel) stays resident in shared memory.
//
// B operand layout: K-major, SWIZZLE_128B: one 32-float K chunk of a row is 128 bytes = one
// swizzle row; 8 rows form a 1024-byte atom (SBO = 1024); the 16-byte column j of row r sits at
// r*128 + ((j ^ (r & 7)) * 16).  One tcgen05.mma consumes K = 8 tf32: the descriptor start
// address advances by 32 bytes, the TMEM A address by 8 columns, per k-step.
#pragma once
#include "common.cuh"
#include "spiral_conv.cuh"

namespace sdvae {
namespace umma {

constexpr int kEpilogueWarps = 4;
constexpr int kSplitSets = 4;
constexpr int kSplitWarps = 4 * kSplitSets;
constexpr int kLoadWarps = 7;
constexpr int kFirstEpilogueWarp = 0;
constexpr int kFirstSplitWarp = kFirstEpilogueWarp + kEpilogueWarps;      // 4
constexpr int kFirstLoadWarp = kFirstSplitWarp + kSplitWarps;             // 20
constexpr int kMmaWarp = kFirstLoadWarp + kLoadWarps;                     // 27: the issuer gets the highest warp id
constexpr int kThreads = (kMmaWarp + 1) * 32;                             // 896 -> 72 registers per thread at launch
// setmaxnreg budgets per warpgroup: 128*56 + 512*80 + 256*64 = 64512 = 896*72
#ifndef SDVAE_REGS_EPI
#define SDVAE_REGS_EPI 56
#define SDVAE_REGS_SPLIT 80
#define SDVAE_REGS_LOAD 64
#endif
constexpr int kRegsEpilogue = SDVAE_REGS_EPI, kRegsSplit = SDVAE_REGS_SPLIT, kRegsLoad = SDVAE_REGS_LOAD;
constexpr int kBM = 128;                  // rows per tile (UMMA M)
constexpr int kMaxAStages = 7;            // TMEM A-operand ring depth limit (64 columns each: 32 hi + 32 lo)
constexpr int kTmemCols = 512;            // [0, 4*NT): two accumulators, [4*NT, 512): A ring
constexpr int kMaxRcap = 192;             // staged rows per (tile, slot) the kernel supports
constexpr int kMaxRaw = kLoadWarps;       // raw ring depth limit: loader warp w owns raw stage w
constexpr int kOutRowBytes = 144;         // output staging: 128-byte rows padded to 144 B (conflict-free both ways)
constexpr int kOutStageBytes = kBM * kOutRowBytes;
constexpr uint32_t kSuspendHintNs = 100000;          // mbarrier.try_wait suspend-time hint
constexpr int kSpinLimit = 1 << 22;                 // failed try_waits before a stuck wait traps
#ifndef SDVAE_BACKOFF_NS
#define SDVAE_BACKOFF_NS 100
#endif
constexpr unsigned kBackoffNs = SDVAE_BACKOFF_NS;    // sleep between polls of the relaxed waits

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per WARP (after a warp-level sync) instead of one per thread.  A sleeping waiter is woken by
// every arrival on the CTA's barriers; with per-thread arrivals (128 per hand-off) the waiting warps of the
// weight-gradient kernel executed 40 wake-up / re-check rounds per wait -- half of all issued instructions
// (profiles/r01_bw_umma_stalls.txt).
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or
// the hint expires) instead of re-polling -- with 21 warps per CTA a polling loop would take issue
// slots from the warps that have work.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) __trap();
    }
}
// The same for the roles that have slack (loaders, splitters, epilogue): a failed try_wait comes back after
// ~40 clk whatever the suspend hint says, and with 20+ warps of a CTA polling, the re-check rounds took
// half of the SM's issue slots from the warps that had work (ncu: 199 M instructions for 8.6 k tiles, 150 M
// of them in wait loops).  Back off between polls; the MMA warp keeps the tight loop.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    int spins = 0;
    do {
        __nanosleep(kBackoffNs);
        if (++spins > kSpinLimit) __trap();
    } while (!mbar_try_wait(bar, parity));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        default: cp_async_wait<6>(); break;
    }
}

// One lane of a converged warp (elect.sync): ptxas then emits the tcgen05 issue sequence as straight
// uniform-datapath code; guarding it with `lane == 0` instead wraps every UTCHMMA in an ELECT retry loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
    return pred != 0;
}

#ifdef SDVAE_NO_SETMAXNREG
template <int N> __device__ __forceinline__ void reg_inc() {}
template <int N> __device__ __forceinline__ void reg_dec() {}
#else
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
#endif

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, issued by one thread for the CTA.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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
// thread t of the warp -> lane (base_lane + t), 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr),
          "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
          "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
          "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- tuning instrumentation, compiled in with -DSDVAE_TUNING only (it costs 5-8 % when present) -------
// SDVAE_DBG bits: 1 no copies, 2 no staged-row reads, 4 no MMAs, 8 no epilogue math/stores, 16 no split/STTM,
// 32 per-role wait-cycle counters of CTA 0 (sdvae_debug_read_prof), 64/128/256 MMA-thread micro-ablations.
#ifdef SDVAE_TUNING
constexpr bool kTuning = true;
#else
constexpr bool kTuning = false;
#endif
#define SDVAE_DBG_ON(args, bit) (::sdvae::umma::kTuning && ((args).dbg & (bit)) != 0)
__device__ long long g_prof[64];
struct WaitClock {
    long long acc; bool on;
    __device__ __forceinline__ WaitClock(bool on_) : acc(0), on(on_) {}
    template <class F> __device__ __forceinline__ void timed(F&& f) {
        if (on) { const long long t0 = clock64(); f(); acc += clock64() - t0; } else f();
    }
};

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

// ---- loader: packed plan + cp.async issue ------------------------------------------------------
// plan_src is PACKED: per (tile, slot) rcap/2 32-bit words (rcap a multiple of 32), two 16-bit source
// rows per word, in the order the loader lanes consume them: lane (rsub = lane >> 3, q = lane & 7) copies
// the 16-byte piece q of the staged rows e = 32*j + 4*t + rsub (t = 0..7) of every group j of 32 rows, and
// finds those eight rows in the four consecutive words 16*j + 4*rsub + (t >> 1) (low half = even t): one
// 16-byte load per lane and group, no shuffles, no per-lane predicates (entries past the count are row 0,
// a valid address; their staged rows are never read).  Per cp.async that leaves: extract, one IMAD.WIDE,
// LDGSTS -- the shuffle-based loader spent ~9 instructions per copy and bounded the kernel.
template <int PV>
struct PlanRegs { uint4 w[PV]; int n; };

template <int PV>
__device__ __forceinline__ void plan_fetch(PlanRegs<PV>& p, const int* plan_cnt, const int* plan_src,
                                           int jt, int S, int s, int rcap, int rsub) {
    p.n = __ldg(plan_cnt + jt * S + s);
    const uint4* src = reinterpret_cast<const uint4*>(plan_src + ((size_t)jt * S + s) * (rcap >> 1)) + rsub;
#pragma unroll
    for (int j = 0; j < PV; ++j) p.w[j] = (32 * j < rcap) ? __ldg(src + 4 * j) : make_uint4(0u, 0u, 0u, 0u);
}

// dst0/dst1: shared address of staged row rsub, swizzled 16-byte column of the even / odd t rows;
// base: global address of piece q of row 0; row_bytes: bytes per source row
template <int PV>
__device__ __forceinline__ void plan_issue(const PlanRegs<PV>& p, uint32_t dst0, uint32_t dst1,
                                           const float* base, uint32_t row_bytes) {
    const char* gb = reinterpret_cast<const char*>(base);
#pragma unroll
    for (int j = 0; j < PV; ++j) {
        if (32 * j < p.n) {
            const uint32_t w[4] = {p.w[j].x, p.w[j].y, p.w[j].z, p.w[j].w};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t row = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                             ::"r"(((t & 1) ? dst1 : dst0) + (uint32_t)(32 * j + 4 * t) * 128u),
                               "l"(gb + (size_t)row * row_bytes));
            }
        }
    }
}

// K permutation of the tile-staged kernels (spiral_conv_tile.cuh): channel held by A column kk of a 32-wide K
// chunk.  Thread q = (kk >> 1) & 3 of a row reads the 16-byte pieces q (channels 4q..4q+3 -> column groups
// n = kk >> 3 = 0, 1) and q + 4 (channels 16+4q.. -> n = 2, 3) of the staged row.
__host__ __device__ constexpr int kperm(int kk) {
    return ((kk >> 4) << 4) + (((kk >> 1) & 3) << 2) + (((kk >> 3) & 1) << 1) + (kk & 1);
}

// ---- weight image ----------------------------------------------------------------------------
// img[chunk][j][32]:  j < NT -> hi part of W row j,  j >= NT -> lo part of row j-NT; rows >= n_real
// are zero.  `transposed` selects the backward-to-input weight  Wt[c, s*Cout + o] = W[o, s*Cin + c]
// read straight from the forward weight (so no separate transpose pass is needed):
//   forward   : n = output channel, k = s*KS + c      -> W[n*ldw + k]            (KS = Cin)
//   transposed: n = input channel c, k = s*KS + o     -> W[o*ldw + s*cin + n]    (KS = Cout)
// `transposed` is a flag word: bit 0 = transposed, bit 1 = K position kk of every 32-wide chunk holds channel
// kperm(kk) (the images of the tile-staged kernels).
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
        const int k = ch * 32 + ((a.transposed & 2) ? kperm(kk) : kk);
        float w = 0.f;
        if (n < a.n_real) {
            if (!(a.transposed & 1)) {
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

// Batched form: blockIdx.y selects the image; entries live in device memory (layout = sdvae_pack_entry).
struct PackEntry { const float* W; float* img; int S, Cin, Cout, transposed, n0, n_cnt; };

__device__ __forceinline__ int pack_tile_n(int N) { return N <= 16 ? 16 : (N <= 32 ? 32 : 64); }

__global__ void umma_pack_weights_batch_kernel(const PackEntry* __restrict__ entries) {
    const PackEntry e = entries[blockIdx.y];
    PackArgs a;
    a.KS = (e.transposed & 1) ? e.Cout : e.Cin;
    a.W = (e.transposed & 1) ? e.W + e.n0 : e.W + (size_t)e.n0 * e.S * e.Cin;
    a.img = e.img; a.NT = pack_tile_n(e.n_cnt); a.S = e.S; a.n_real = e.n_cnt; a.ldw = e.S * e.Cin;
    a.transposed = e.transposed; a.cin = e.Cin;
    const int K = a.S * a.KS;
    const int total = (K / 32) * 2 * a.NT * 32;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int kk = t & 31;
        const int j = (t >> 5) % (2 * a.NT);
        const int ch = (t >> 5) / (2 * a.NT);
        const int n = j % a.NT, part = j / a.NT;
        const int k = ch * 32 + ((a.transposed & 2) ? kperm(kk) : kk);
        float w = 0.f;
        if (n < a.n_real) {
            if (!(a.transposed & 1)) {
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
struct UmmaArgs {
    const float* in;          // [B, in_rows, KS]
    const int* plan_cnt;      // [L, S]            staged rows of (tile, slot)
    const int* plan_src;      // [L, S, rcap/2]    packed source rows of the staged rows (see plan_fetch)
    const int* plan_cell;     // [L, S, 128]       start | count << 16 : staged rows summed into tile row lr
    const float* wimg;        // packed weight image
    const float* bias;        // [n_real] or nullptr
    const float* gate;        // EPI_GATE: out *= elu'(gate), aligned with out
    float* out;               // [B, out_rows, ldo]
    int B, in_rows, out_rows, L, S, rcap;
    int n_real, ldo, epi;
    int nraw;                 // raw ring depth (= active loader warps)
    int ostage;               // 1: epilogue stages output rows in shared memory (coalesced stores); NT == 32 only
    int dbg;                  // ablation switches for tuning runs (SDVAE_DBG): 1 no copies, 2 no staged-row reads, 4 no MMAs
};


// Position of a role in the CTA's static schedule: chunk g = (tile iteration, chunk ch of the tile),
// kept incrementally (no integer division in the per-chunk loops).  `step` chunks per advance.
struct ChunkCursor {
    int g, ch, b, jt, rs, as;
    uint32_t rph, aph;
    int nch, nraw, nast, L, db, djt;
    __device__ __forceinline__ ChunkCursor(const UmmaArgs& a, int nch_, int nraw_, int nast_)
        : g(0), ch(0), rs(0), as(0), rph(0), aph(0), nch(nch_), nraw(nraw_), nast(nast_), L(a.L) {
        b = (int)blockIdx.x / L; jt = (int)blockIdx.x - b * L;
        db = (int)gridDim.x / L; djt = (int)gridDim.x - db * L;       // tile stride, split into (mesh, tile-in-mesh)
    }
    __device__ __forceinline__ void advance(int step) {
        g += step;
        rs += step; while (rs >= nraw) { rs -= nraw; rph ^= 1; }
        as += step; while (as >= nast) { as -= nast; aph ^= 1; }
        ch += step;
        while (ch >= nch) {
            ch -= nch; b += db; jt += djt;
            if (jt >= L) { jt -= L; ++b; }
        }
    }
};

template <int KS, int NT>
struct UmmaCfg {
    static constexpr int CPS = KS / 32;                  // 32-wide chunks per spiral slot
    static constexpr int B_CHUNK = 2 * NT * 128;
    static constexpr int ACC_COLS = 4 * NT;              // two accumulators of 2*NT columns
    static constexpr int MAX_AST = (kTmemCols - ACC_COLS) / 64 < kMaxAStages ? (kTmemCols - ACC_COLS) / 64 : kMaxAStages;
    static size_t b_bytes(int S) { return (size_t)S * CPS * B_CHUNK; }
    static int raw_stages(int S, int rcap) {
        const long long budget = 226LL * 1024 - 1024 /*align*/ - 1024 /*barriers*/ - (long long)b_bytes(S);
        long long st = budget / ((long long)rcap * 128);
        return (int)(st > kMaxRaw ? kMaxRaw : st);
    }
    // the output staging tile is used when it fits WITHOUT shortening the raw ring
    static bool out_stage(int S, int rcap) {
        if (NT != 32) return false;
        const long long budget = 226LL * 1024 - 2048 - (long long)b_bytes(S) - kOutStageBytes;
        return budget / ((long long)rcap * 128) >= kMaxRaw;
    }
    static size_t smem_bytes(int S, int rcap, int nraw, bool ostage) {
        return 1024 + b_bytes(S) + (size_t)nraw * rcap * 128 + (ostage ? kOutStageBytes : 0) + 1024;
    }
};

// UNIFORM: forward tile plan (staged row e of a (tile, slot) IS tile row e): the splitter needs no cell table
template <int KS, int NT, bool UNIFORM>
__global__ void __launch_bounds__(kThreads, 1)
gc_umma_kernel(const UmmaArgs a) {
    using Cfg = UmmaCfg<KS, NT>;
    constexpr int CPS = Cfg::CPS, B_CHUNK = Cfg::B_CHUNK, ACOL = Cfg::ACC_COLS;
    const int S = a.S;
    const int NCH = S * CPS;
    const int NRAW = a.nraw;
    // TMEM A-ring depth and the number of active splitter sets.  The parity waits need every waiter within
    // one phase of its barrier:
    //  * a set handles every NS-th chunk, so before it waits for the MMAs of chunk g - NAST (a_empty) it has
    //    seen those of chunk g - NS - NAST complete; NS <= NAST keeps that within one phase;
    //  * once the MMAs of chunk g - NAST are done every chunk up to g - NAST has been loaded; NAST <= NRAW
    //    then keeps the raw barrier of chunk g's stage at most one phase behind its waiter.
    const int NAST = NRAW < Cfg::MAX_AST ? NRAW : Cfg::MAX_AST;
    const int NS = NAST < kSplitSets ? NAST : kSplitSets;
    const int RAW_STAGE = a.rcap * 128;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* B_s = smem;                                   // [NCH][2NT][128 B]   resident weight image
    uint8_t* R_s = B_s + (size_t)NCH * B_CHUNK;            // [NRAW][rcap][128 B] staged source rows (swizzled)
    uint8_t* O_s = R_s + (size_t)NRAW * RAW_STAGE;         // [128][144 B] output staging (a.ostage only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(O_s + (a.ostage ? kOutStageBytes : 0));
    uint64_t* raw_full = bars;                             // [NRAW]  loaders   -> splitters
    uint64_t* raw_empty = bars + kMaxRaw;                  // [NRAW]  splitters -> loaders
    uint64_t* a_full = bars + 2 * kMaxRaw;                 // [NAST]  splitters -> MMA
    uint64_t* a_empty = a_full + kMaxAStages;              // [NAST]  MMA (commit) -> splitters
    uint64_t* t_full = a_empty + kMaxAStages;              // [2]     MMA (commit) -> epilogue
    uint64_t* t_empty = t_full + 2;                        // [2]     epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < NRAW; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, 4); }
        for (int i = 0; i < kMaxAStages; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, kEpilogueWarps); }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    {   // resident weight image (generic-proxy writes, read by the MMA through the async proxy)
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
    const int G = my_tiles * NCH;                           // chunks this CTA processes

    if (warp >= kFirstLoadWarp) {
        reg_dec<kRegsLoad>();
        if (warp == kMmaWarp) {
            // ================= MMA issuer =================
            // ONE elected thread walks the schedule.  Its per-chunk instruction stream is the serial spine of
            // the kernel (with every other role ablated the kernel still took 324 clk per chunk, all of it this
            // loop: ~70 mostly uniform-datapath instructions per chunk, each waiting for the previous one), so:
            // no per-chunk warp sync, shared-memory descriptors by one 32-bit add from a precomputed base, and
            // the readiness of the NEXT chunk is polled before the MMAs of the current one are issued, which
            // hides the ~90 clk try_wait round trip behind them.
            if (elect_one()) {
                constexpr uint32_t IDESC1 = idesc_tf32(kBM, 2 * NT);
                constexpr uint32_t IDESC2 = idesc_tf32(kBM, NT);
                const uint64_t desc0 = smem_desc_sw128(smem_u32(B_s));
                const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
                const bool no_mma = SDVAE_DBG_ON(a, 4);
                int as = 0; uint32_t aph = 0;
                const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0;
                WaitClock w_afull(prof), w_tempty(prof);
                const long long t_begin = prof ? clock64() : 0;
                bool ready = my_tiles > 0 && mbar_try_wait(a_full, 0u);
#pragma unroll 1
                for (int it = 0; it < my_tiles; ++it) {
                    const int acc = it & 1;
                    w_tempty.timed([&] { mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1); });
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * NT);
#pragma unroll 1
                    for (int ch = 0; ch < NCH; ++ch) {
                        if (!ready) w_afull.timed([&] { mbar_wait(a_full + as, aph); });
                        if (!SDVAE_DBG_ON(a, 128)) tc_fence_after();
                        const uint32_t a_hi = tmem_base + (uint32_t)(ACOL + as * 64), a_lo = a_hi + 32;
                        const uint32_t dl = desc_lo0 + (uint32_t)(ch * (B_CHUNK >> 4));
                        uint64_t* const my_empty = a_empty + as;
                        if (++as == NAST) { as = 0; aph ^= 1; }
                        ready = !SDVAE_DBG_ON(a, 256) && (ch + 1 < NCH || it + 1 < my_tiles) && mbar_try_wait(a_full + as, aph);
                        if (!no_mma) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(dl + 2u * k);
                                umma_tf32_ts(d_tmem, a_hi + k * 8, bd, IDESC1, (ch | k) != 0);
                                umma_tf32_ts(d_tmem, a_lo + k * 8, bd, IDESC2, 1u);
                            }
                        }
                        if SDVAE_DBG_ON(a, 64) { mbar_arrive(my_empty); if (ch == NCH - 1) mbar_arrive(t_full + acc); }
                        else { umma_commit(my_empty); if (ch == NCH - 1) umma_commit(t_full + acc); }
                    }
                }
                if (prof) { g_prof[0] = clock64() - t_begin; g_prof[1] = w_afull.acc; g_prof[2] = w_tempty.acc; g_prof[3] = G; }
            }
            __syncwarp();
        } else {
            // ================= loaders =================
            // loader warp w (< NRAW) owns raw stage w and takes the chunks g = w (mod NRAW) whole: 4 staged rows
            // per cp.async instruction (8 lanes x 16 B per row piece), the per-chunk bookkeeping is paid by one
            // warp, and the warps together keep NRAW chunks in flight.  (One stage per warp also keeps every
            // waiter within one phase of its mbarrier, which the parity wait requires.)
            const int lw = warp - kFirstLoadWarp;
            const int q = lane & 7, rsub = lane >> 3;
            const uint32_t sw0 = (uint32_t)((q ^ rsub) << 4), sw1 = (uint32_t)((q ^ (rsub + 4)) << 4);
            const uint32_t raw_base = smem_u32(R_s) + (uint32_t)rsub * 128u;
            ChunkCursor cur(a, NCH, NRAW, NAST);
            cur.advance(lw);
            if (lw >= NRAW) cur.g = G;                                // more loader warps than stages: idle
            // Forward plans stage <= 128 rows per chunk: their plan words are fetched one own-chunk ahead, so
            // the issue loop never waits for them.  Backward plans (up to kMaxRcap rows) fetch at chunk start.
            constexpr int PV = UNIFORM ? 4 : kMaxRcap / 32;
            const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0 && lw == 0;
            WaitClock w_rempty(prof), w_copy(prof), w_plan(prof);
            const long long t_begin = prof ? clock64() : 0;
            PlanRegs<PV> nxt;
            if (UNIFORM && cur.g < G) plan_fetch(nxt, a.plan_cnt, a.plan_src, cur.jt, S, cur.ch / CPS, a.rcap, rsub);
#pragma unroll 1
            while (cur.g < G) {
                PlanRegs<PV> now;
                w_plan.timed([&] {
                    if (UNIFORM) now = nxt;
                    else plan_fetch(now, a.plan_cnt, a.plan_src, cur.jt, S, cur.ch / CPS, a.rcap, rsub);
                    if (prof && now.n < 0) g_prof[63] = now.w[0].x;      // force the loads to have landed
                });
                const int rs = cur.rs, h = cur.ch % CPS;
                const uint32_t rph = cur.rph;
                const float* base = a.in + (size_t)cur.b * a.in_rows * KS + h * 32 + 4 * q;
                cur.advance(NRAW);
                if (UNIFORM && cur.g < G) plan_fetch(nxt, a.plan_cnt, a.plan_src, cur.jt, S, cur.ch / CPS, a.rcap, rsub);
                w_rempty.timed([&] { mbar_wait_relaxed(raw_empty + rs, rph ^ 1); });
                const uint32_t dst = raw_base + (uint32_t)rs * (uint32_t)RAW_STAGE;
                if (!SDVAE_DBG_ON(a, 1)) plan_issue(now, dst + sw0, dst + sw1, base, KS * 4u);
                cp_async_commit();
                w_copy.timed([&] { cp_async_wait<0>(); });
                warp_arrive(raw_full + rs, lane);
            }
            if (prof && lane == 0) { g_prof[8] = clock64() - t_begin; g_prof[9] = w_rempty.acc; g_prof[10] = w_copy.acc; g_prof[11] = w_plan.acc; }
        }
    } else if (warp < kFirstSplitWarp) {
        reg_dec<kRegsEpilogue>();
        // ================= epilogue =================
        const int q4 = warp & 3;                                  // TMEM lane quarter this warp may access
        const int EPI = a.epi;
        const int n_real = a.n_real, ldo = a.ldo;
        const bool has_bias = (EPI == EPI_BIAS || EPI == EPI_BIAS_ELU) && a.bias != nullptr;
        // NT >= 32: full-width rows, 16-byte accesses (the host checks n_real == NT and the alignments);
        // NT == 16: narrow outputs (e.g. the 3-channel output layer), scalar tail
        constexpr bool vec_ok = NT >= 32;
        // output staging is compiled into the forward (uniform) instantiations only: the backward plans stage up to
        // 192 rows per chunk and leave no room for it, and the extra code cost the ragged kernel 10 %
        constexpr bool kCanStage = NT == 32 && UNIFORM;
        int b = (int)blockIdx.x / a.L, jt = (int)blockIdx.x - b * a.L;
        const int db = (int)gridDim.x / a.L, djt = (int)gridDim.x - db * a.L;
        const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0 && warp == 0;
        WaitClock w_tfull(prof);
        const long long t_begin = prof ? clock64() : 0;
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            w_tfull.timed([&] { mbar_wait_relaxed(t_full + acc, (it >> 1) & 1); });
            tc_fence_after();
            const int r = jt * kBM + q4 * 32 + lane;              // row inside the mesh
            const bool row_ok = r < a.out_rows;
            const size_t m = (size_t)b * a.out_rows + r;
            const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * 2 * NT);
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 16) {
                float v[16], d2[16];
                tmem_ld16(t_row + c0, v);
                tmem_ld16(t_row + NT + c0, d2);
                tmem_ld_wait();
                if (c0 + 16 >= NT) {            // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    warp_arrive(t_empty + acc, lane);
                }
                if (!row_ok || SDVAE_DBG_ON(a, 8)) continue;      // (rows past the mesh: nothing staged, nothing stored)
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += d2[j];
                float* orow = a.out + m * ldo + c0;
                const float* grow = a.gate + m * ldo + c0;
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
                    if (EPI == EPI_GATE) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 gt = ldg4(grow + j);
                            v[j] *= elu_grad_from_out(gt.x); v[j + 1] *= elu_grad_from_out(gt.y);
                            v[j + 2] *= elu_grad_from_out(gt.z); v[j + 3] *= elu_grad_from_out(gt.w);
                        }
                    }
                    if (kCanStage && a.ostage) {
                        // row-per-thread 16-byte global stores touch 32 lines per instruction; stage the row in
                        // shared memory instead and store 4 whole rows per instruction below
                        float4* srow = reinterpret_cast<float4*>(O_s + (q4 * 32 + lane) * kOutRowBytes + c0 * 4);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) srow[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                } else {
                    // narrow / unaligned outputs (e.g. the 3-channel output layer): scalar tail
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (c0 + j < n_real) {
                            float o = v[j];
                            if (EPI == EPI_GATE) o *= elu_grad_from_out(__ldg(grow + j));
                            orow[j] = o;
                        }
                    }
                }
            }
            if (kCanStage && a.ostage && !SDVAE_DBG_ON(a, 8)) {
                // coalesced write-out of this warp's 32 rows: lane -> (row 4k + lane/8, 16-byte piece lane%8)
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
        if (prof && lane == 0) { g_prof[16] = clock64() - t_begin; g_prof[17] = w_tfull.acc; }
    } else {
        reg_inc<kRegsSplit>();
        // ================= splitters =================
        const int set = (warp - kFirstSplitWarp) >> 2;            // chunks g = set (mod NS)
        const int q4 = warp & 3;                                  // TMEM lane quarter
        const int lr = q4 * 32 + lane;                            // tile row owned by this thread
        const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)ACOL;
        if (set < NS) {
            ChunkCursor cur(a, NCH, NRAW, NAST);                  // chunk being processed
            cur.advance(set);
            const bool prof = SDVAE_DBG_ON(a, 32) && blockIdx.x == 0 && warp == kFirstSplitWarp;
            WaitClock w_aempty(prof), w_rfull(prof), w_st(prof);
            const long long t_begin = prof ? clock64() : 0;
            uint32_t cell = 0u;
            if (!UNIFORM && cur.g < G)
                cell = (uint32_t)__ldg(a.plan_cell + ((size_t)cur.jt * S + cur.ch / CPS) * kBM + lr);
#pragma unroll 1
            while (cur.g < G) {
                const int rs = cur.rs, as = cur.as;
                const uint32_t rph = cur.rph, aph = cur.aph;
                cur.advance(NS);
                uint32_t cell_next = 0u;                          // one own-chunk ahead
                if (!UNIFORM && cur.g < G)
                    cell_next = (uint32_t)__ldg(a.plan_cell + ((size_t)cur.jt * S + cur.ch / CPS) * kBM + lr);
                // Order matters (see NAST above): a_empty first, then the raw stage.
                w_aempty.timed([&] { mbar_wait_relaxed(a_empty + as, aph ^ 1); });
                w_rfull.timed([&] { mbar_wait_relaxed(raw_full + rs, rph); });
                const uint8_t* stage = R_s + (size_t)rs * RAW_STAGE;
                float v[32];
                if SDVAE_DBG_ON(a, 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (float)(lr + j);
                } else if (UNIFORM) {
                    // rows past the end of the mesh were never staged: their A rows are garbage, and so are the
                    // matching accumulator rows, which the epilogue never stores (MMA rows are independent)
                    // (a stage holds rcap rows: small tables have rcap < 128, and rows >= rcap lie outside it)
                    const uint8_t* row = stage + (lr < a.rcap ? lr : 0) * 128;
                    const int x7 = lr & 7;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = *reinterpret_cast<const float4*>(row + ((j ^ x7) << 4));
                        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
                    }
                } else {
                    const int e0 = (int)(cell & 0xffffu), cnt = (int)(cell >> 16);
                    {   // first row of the cell: plain copy (most cells hold exactly one row); empty cells read
                        // staged row 0 and keep zeros
                        const uint8_t* row = stage + (cnt > 0 ? e0 : 0) * 128;
                        const int x7 = (cnt > 0 ? e0 : 0) & 7;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = *reinterpret_cast<const float4*>(row + ((j ^ x7) << 4));
                            v[4 * j] = cnt > 0 ? t.x : 0.f; v[4 * j + 1] = cnt > 0 ? t.y : 0.f;
                            v[4 * j + 2] = cnt > 0 ? t.z : 0.f; v[4 * j + 3] = cnt > 0 ? t.w : 0.f;
                        }
                    }
#pragma unroll 1
                    for (int e = e0 + 1; e < e0 + cnt; ++e) {         // in-order sum: deterministic scatter-add
                        const uint8_t* row = stage + e * 128;
                        const int x7 = e & 7;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = *reinterpret_cast<const float4*>(row + ((j ^ x7) << 4));
                            v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                        }
                    }
                }
                if (!SDVAE_DBG_ON(a, 16)) {
                float lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) { float h; split_tf32f(v[j], h, lo[j]); v[j] = h; }
                tc_fence_after();
                const uint32_t t_a = t_lane + (uint32_t)(as * 64);
                tmem_st32(t_a, v);
                tmem_st32(t_a + 32, lo);
                }
                warp_arrive(raw_empty + rs, lane);    // the staged rows are in registers (consumed by the stores above)
                w_st.timed([&] { tmem_st_wait(); });
                tc_fence_before();
                warp_arrive(a_full + as, lane);
                cell = cell_next;
            }
            if (prof && lane == 0) { g_prof[24] = clock64() - t_begin; g_prof[25] = w_aempty.acc; g_prof[26] = w_rfull.acc; g_prof[27] = w_st.acc; }
        }
    }

    // ---- teardown ----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace umma
}  // namespace sdvae
