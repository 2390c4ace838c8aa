// Spiral convolution kernels (reference model.py:27-41 + F.elu at model.py:68,84).
//
//   y[b,v,o] = act( bias[o] + sum_{s<S} sum_{c<Cin} W[o, s*Cin+c] * x[b, idx[v,s], c] )
//
// One tiled "gather-contraction" kernel serves three uses:
//   * forward            : uniform table idx[Vout,S]; A-tile = gathered rows of x
//   * backward-to-input  : ragged table (for each input vertex u and slot s, the
//                          list of output rows r with idx[r,s]==u, ascending); the
//                          A-tile cell (u,s) is the SUM of those dpre rows, done in
//                          a fixed order -> deterministic, atomic-free scatter-add
//   * dense [M,K]x[K,N]  : idx == nullptr (identity gather, S = 1)
// The A-tile (BM rows x 32 k) and the matching 32-wide weight slab are staged in
// shared memory (cp.async double buffering), then contracted with register-tiled
// fp32 FMAs; bias + ELU (forward) or the ELU-derivative gate (backward) run in
// the epilogue, which leaves through shared memory so that global stores are
// coalesced 16-byte rows.
#pragma once
#include "common.cuh"

namespace sdvae {

enum : int { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_ELU = 2, EPI_GATE = 3 };

struct GcArgs {
    const float* in;        // [B, in_rows, KS]
    const int* idx;         // uniform: [Vout, S] source vertex per (row, slot); nullptr = identity
    const int* cell_ptr;    // ragged: [Vout*S + 1]
    const int* cell_src;    // ragged: [E] source rows, ascending inside a cell
    const float* W;         // row n, column k  ->  W[n*ldw + k],  k = s*KS + c
    const float* bias;      // [n_real] or nullptr
    const float* gate;      // EPI_GATE: activation y aligned with out ([M, ldo]); out *= elu'(y)
    float* out;             // [M, ldo]
    long long M;            // B * Vout
    int in_rows;            // vertices per mesh in `in`
    int Vout;
    int S;
    int ldw;
    int ldo;                // row stride of out / gate (= total N)
    int n_real;             // total number of valid output columns
    int epi;                // EPI_*
};

// ---------------------------------------------------------------------------
// Tiled kernel.  KS = channels per slot (3 -> "small K" single-chunk path with
// S*KS <= 32; 32/64 -> 32-wide chunks).  NT = tile width (n_real padded up).
// Lanes form RL x CL; each thread owns TM x TN outputs, rows rl + RL*i and
// columns cl + CL*j (interleaved -> conflict-free LDS.128 with row pad 36).
// ---------------------------------------------------------------------------
template <int KS, int NT, int TM, int TN, int NWARPS>
struct GcCfg {
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int CL = NT / TN;
    static constexpr int RL = 32 / CL;
    static constexpr int WM = RL * TM;
    static constexpr int BM = NWARPS * WM;
    static constexpr int P = 36;
    static constexpr bool SMALLK = KS < 32;
    static constexpr int NBUF = SMALLK ? 1 : 2;
    static constexpr int CPS = SMALLK ? 1 : KS / 32;
    static constexpr int CP = NT + CL;
    static size_t smem_bytes(int S, bool ragged) {
        size_t tiles = (size_t)NBUF * (BM + NT) * P * sizeof(float);
        size_t book = ragged ? (size_t)BM * 2 * sizeof(int) : (size_t)BM * S * sizeof(int);
        size_t epi = (size_t)BM * CP * sizeof(float);
        size_t main_part = tiles + book;
        return main_part > epi ? main_part : epi;
    }
};

template <int KS, int NT, int TM, int TN, int NWARPS, bool RAGGED>
__global__ void __launch_bounds__(NWARPS * 32)
gc_tile_kernel(const GcArgs a) {
    const int EPI = a.epi;
    using Cfg = GcCfg<KS, NT, TM, TN, NWARPS>;
    constexpr int THREADS = Cfg::THREADS, CL = Cfg::CL, RL = Cfg::RL, WM = Cfg::WM, BM = Cfg::BM;
    constexpr int P = Cfg::P, NBUF = Cfg::NBUF, CPS = Cfg::CPS, CP = Cfg::CP;
    constexpr bool SMALLK = Cfg::SMALLK;
    static_assert(32 % CL == 0 && NT % TN == 0, "bad lane layout");

    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;
    float* W_s = A_s + NBUF * BM * P;
    int* I_s = reinterpret_cast<int*>(W_s + NBUF * NT * P);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cl = lane % CL, rl = lane / CL;
    const long long base = (long long)blockIdx.x * BM;
    const int S = a.S;
    const int nblk = blockIdx.y * NT;
    const float* Wg = a.W + (size_t)nblk * a.ldw;

    // ---- per-row bookkeeping ------------------------------------------------
    if (!RAGGED) {
        for (int e = tid; e < BM * S; e += THREADS) {
            const int lr = e / S, s = e - lr * S;
            const long long m = base + lr;
            int v = 0;
            if (m < a.M) {
                const int b = (int)(m / a.Vout);
                const int r = (int)(m - (long long)b * a.Vout);
                v = b * a.in_rows + (a.idx ? __ldg(a.idx + r * S + s) : r);
            }
            I_s[e] = v;
        }
    } else {
        for (int lr = tid; lr < BM; lr += THREADS) {
            const long long m = base + lr;
            int b = 0, r = -1;
            if (m < a.M) {
                b = (int)(m / a.Vout);
                r = (int)(m - (long long)b * a.Vout);
            }
            I_s[2 * lr] = b * a.in_rows;
            I_s[2 * lr + 1] = r;
        }
    }
    __syncthreads();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto compute = [&](int buf, int k4_count) {
        const float* As = A_s + (buf * BM + warp * WM + rl) * P;
        const float* Ws = W_s + (buf * NT + cl) * P;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            if (k4 < k4_count) {
                float4 av[TM], wv[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(As + i * RL * P + 4 * k4);
#pragma unroll
                for (int j = 0; j < TN; ++j) wv[j] = *reinterpret_cast<const float4*>(Ws + j * CL * P + 4 * k4);
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        acc[i][j] = fmaf(av[i].x, wv[j].x, acc[i][j]);
                        acc[i][j] = fmaf(av[i].y, wv[j].y, acc[i][j]);
                        acc[i][j] = fmaf(av[i].z, wv[j].z, acc[i][j]);
                        acc[i][j] = fmaf(av[i].w, wv[j].w, acc[i][j]);
                    }
            }
        }
    };

    if constexpr (SMALLK) {
        // K = S*KS <= 32 : one chunk, scalar gather (rows of x are KS*4 bytes, unaligned for 16B)
        const int K = S * KS;
        for (int e = tid; e < BM * 32; e += THREADS) {
            const int lr = e >> 5, k = e & 31;
            float v = 0.f;
            if (k < K) {
                const int s = k / KS, c = k - s * KS;
                if (!RAGGED) {
                    v = __ldg(a.in + (size_t)I_s[lr * S + s] * KS + c);
                } else {
                    const int r = I_s[2 * lr + 1];
                    if (r >= 0) {
                        const int cell = r * S + s;
                        const int e0 = __ldg(a.cell_ptr + cell), e1 = __ldg(a.cell_ptr + cell + 1);
                        for (int q = e0; q < e1; ++q)
                            v += __ldg(a.in + (size_t)(I_s[2 * lr] + __ldg(a.cell_src + q)) * KS + c);
                    }
                }
            }
            A_s[lr * P + k] = v;
        }
        for (int e = tid; e < NT * 32; e += THREADS) {
            const int n = e >> 5, k = e & 31;
            W_s[n * P + k] = (k < K && nblk + n < a.n_real) ? __ldg(Wg + (size_t)n * a.ldw + k) : 0.f;
        }
        __syncthreads();
        compute(0, (K + 3) >> 2);
        __syncthreads();
    } else {
        constexpr int NCELL = BM * 8 / THREADS;   // 16-byte cells per thread per chunk
        static_assert((BM * 8) % THREADS == 0, "tile/threads mismatch");
        const int NCH = S * CPS;

        auto w_fill = [&](int ch, int buf) {
            for (int e = tid; e < NT * 8; e += THREADS) {
                const int n = e >> 3, q = e & 7;
                float* dst = W_s + (buf * NT + n) * P + 4 * q;
                if (nblk + n < a.n_real) cp_async16(dst, Wg + (size_t)n * a.ldw + ch * 32 + 4 * q);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto a_fill_async = [&](int ch, int buf) {
            const int s = ch / CPS, h = ch - s * CPS;
#pragma unroll
            for (int i = 0; i < NCELL; ++i) {
                const int c = tid + i * THREADS;
                const int lr = c >> 3, q = c & 7;
                cp_async16(A_s + (buf * BM + lr) * P + 4 * q,
                           a.in + (size_t)I_s[lr * S + s] * KS + h * 32 + 4 * q);
            }
        };
        auto a_load_ragged = [&](int ch, float4 (&r4)[NCELL]) {
            const int s = ch / CPS, h = ch - s * CPS;
#pragma unroll
            for (int i = 0; i < NCELL; ++i) {
                const int c = tid + i * THREADS;
                const int lr = c >> 3, q = c & 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int r = I_s[2 * lr + 1];
                if (r >= 0) {
                    const int cell = r * S + s;
                    const int e0 = __ldg(a.cell_ptr + cell), e1 = __ldg(a.cell_ptr + cell + 1);
                    const float* rowbase = a.in + (size_t)I_s[2 * lr] * KS + h * 32 + 4 * q;
                    for (int e = e0; e < e1; ++e) {
                        const float4 t = ldg4(rowbase + (size_t)__ldg(a.cell_src + e) * KS);
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                    }
                }
                r4[i] = v;
            }
        };
        auto a_store_ragged = [&](int buf, const float4 (&r4)[NCELL]) {
#pragma unroll
            for (int i = 0; i < NCELL; ++i) {
                const int c = tid + i * THREADS;
                const int lr = c >> 3, q = c & 7;
                *reinterpret_cast<float4*>(A_s + (buf * BM + lr) * P + 4 * q) = r4[i];
            }
        };

        float4 stage[RAGGED ? NCELL : 1];
        if constexpr (RAGGED) {
            a_load_ragged(0, stage);
            w_fill(0, 0);
            cp_async_commit();
            a_store_ragged(0, stage);
        } else {
            a_fill_async(0, 0);
            w_fill(0, 0);
            cp_async_commit();
        }
        for (int ch = 0; ch < NCH; ++ch) {
            const int buf = ch & 1;
            const bool more = ch + 1 < NCH;
            if (more) {
                if constexpr (RAGGED) a_load_ragged(ch + 1, stage);
                else a_fill_async(ch + 1, buf ^ 1);
                w_fill(ch + 1, buf ^ 1);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            compute(buf, 8);
            if constexpr (RAGGED) {
                if (more) a_store_ragged(buf ^ 1, stage);
            }
            __syncthreads();
        }
    }

    // ---- epilogue: registers -> smem (conflict-free) -> coalesced global ------
    float* C_s = smem;
    {
        float bj[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = nblk + cl + CL * j;
            bj[j] = ((EPI == EPI_BIAS || EPI == EPI_BIAS_ELU) && a.bias && col < a.n_real) ? __ldg(a.bias + col) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                float v = acc[i][j] + bj[j];
                if (EPI == EPI_BIAS_ELU) v = elu_fast(v);
                C_s[(warp * WM + rl + RL * i) * CP + cl + CL * j] = v;
            }
    }
    __syncthreads();
    const int ncols = min(NT, a.n_real - nblk);
    if ((NT % 4 == 0) && ncols == NT && (a.ldo & 3) == 0) {
        constexpr int Q = NT / 4;
        for (int e = tid; e < BM * Q; e += THREADS) {
            const int row = e / Q, c4 = e - row * Q;
            const long long m = base + row;
            if (m < a.M) {
                float4 v = *reinterpret_cast<const float4*>(C_s + row * CP + 4 * c4);
                const size_t off = (size_t)m * a.ldo + nblk + 4 * c4;
                if (EPI == EPI_GATE) {
                    const float4 g = ldg4(a.gate + off);
                    v.x *= elu_grad_from_out(g.x); v.y *= elu_grad_from_out(g.y);
                    v.z *= elu_grad_from_out(g.z); v.w *= elu_grad_from_out(g.w);
                }
                *reinterpret_cast<float4*>(a.out + off) = v;
            }
        }
    } else {
        for (int e = tid; e < BM * ncols; e += THREADS) {
            const int row = e / ncols, c = e - row * ncols;
            const long long m = base + row;
            if (m < a.M) {
                float v = C_s[row * CP + c];
                const size_t off = (size_t)m * a.ldo + nblk + c;
                if (EPI == EPI_GATE) v *= elu_grad_from_out(__ldg(a.gate + off));
                a.out[off] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Generic fallback: any KS / N / S, one thread per output element.  Slow but
// shape-agnostic; also the on-device cross-check for the tiled kernel.
// ---------------------------------------------------------------------------
template <bool RAGGED>
__global__ void gc_generic_kernel(const GcArgs a, int KS, int epi) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = a.M * a.n_real;
    if (t >= total) return;
    const long long m = t / a.n_real;
    const int n = (int)(t - m * a.n_real);
    const int b = (int)(m / a.Vout);
    const int r = (int)(m - (long long)b * a.Vout);
    const float* wrow = a.W + (size_t)n * a.ldw;
    float acc = 0.f;
    for (int s = 0; s < a.S; ++s) {
        if (!RAGGED) {
            const int src = a.idx ? a.idx[r * a.S + s] : r;
            const float* xr = a.in + ((size_t)b * a.in_rows + src) * KS;
            for (int c = 0; c < KS; ++c) acc = fmaf(__ldg(xr + c), __ldg(wrow + s * KS + c), acc);
        } else {
            const int cell = r * a.S + s;
            const int e0 = a.cell_ptr[cell], e1 = a.cell_ptr[cell + 1];
            for (int c = 0; c < KS; ++c) {
                float xs = 0.f;
                for (int e = e0; e < e1; ++e)
                    xs += __ldg(a.in + ((size_t)b * a.in_rows + a.cell_src[e]) * KS + c);
                acc = fmaf(xs, __ldg(wrow + s * KS + c), acc);
            }
        }
    }
    if ((epi == EPI_BIAS || epi == EPI_BIAS_ELU) && a.bias) acc += a.bias[n];
    if (epi == EPI_BIAS_ELU) acc = elu_f(acc);
    const size_t off = (size_t)m * a.ldo + n;
    if (epi == EPI_GATE) acc *= elu_grad_from_out(a.gate[off]);
    a.out[off] = acc;
}

// ---------------------------------------------------------------------------
// Weight gradient:  dW[n, s*KS+c] = sum_m g[m,n] * x[b(m), idx[r(m),s], c],
//                   db[n]         = sum_m g[m,n]
// Each CTA streams a contiguous range of rows m through shared memory (gathered
// x rows re-built on the fly, never materialised in HBM) and keeps a TK x TN
// register block of the [K x N] outer-product sum per thread.  CTAs write
// partial sums; a second kernel adds them in CTA order -> run-to-run identical.
// ---------------------------------------------------------------------------
struct BwArgs {
    const float* in;        // [B, in_rows, KS]
    const int* idx;         // [Vout, S]
    const float* g;         // [M, n_real]  (already multiplied by elu')
    float* part;            // [nsplit, n_real, K]
    float* part_b;          // [nsplit, n_real]
    long long M;
    long long rows_per_cta; // multiple of BMW
    int in_rows, Vout, n_real;
};

template <int KS, int S_, int NT, int TK, int TN, int BMW>
struct BwCfg {
    static constexpr int K = KS * S_;
    static constexpr int KP = (K + TK - 1) / TK * TK;
    static constexpr int KROW = (KP + 3) / 4 * 4;
    static constexpr int NKG = KP / TK;
    static constexpr int NNG = NT / TN;
    static constexpr int THREADS = NKG * NNG;
    static constexpr size_t SMEM = (size_t)2 * BMW * (KROW + NT) * sizeof(float);
};

template <int KS, int S_, int NT, int TK, int TN, int BMW>
__global__ void __launch_bounds__(BwCfg<KS, S_, NT, TK, TN, BMW>::THREADS)
bw_outer_kernel(const BwArgs a) {
    using Cfg = BwCfg<KS, S_, NT, TK, TN, BMW>;
    constexpr int K = Cfg::K, KP = Cfg::KP, KROW = Cfg::KROW, NKG = Cfg::NKG, THREADS = Cfg::THREADS;
    constexpr bool VEC = (KS % 4 == 0);
    static_assert(TK == 2 || TK == 4 || TK == 8, "TK");
    static_assert(TN == 4 || TN == 8, "TN");

    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;                         // [2][BMW][KROW]
    float* G_s = smem + 2 * BMW * KROW;        // [2][BMW][NT]

    const int tid = threadIdx.x;
    const int kg = tid % NKG, ng = tid / NKG;
    const long long m_begin = (long long)blockIdx.x * a.rows_per_cta;
    const long long m_end = min(a.M, m_begin + a.rows_per_cta);

    // zero the padding columns once (cp.async never touches them)
    for (int e = tid; e < 2 * BMW * KROW; e += THREADS) if ((e % KROW) >= K) A_s[e] = 0.f;
    for (int e = tid; e < 2 * BMW * NT; e += THREADS) if ((e % NT) >= a.n_real) G_s[e] = 0.f;
    __syncthreads();

    auto fill = [&](long long m0, int buf) {
        const int b0 = (int)(m0 / a.Vout);
        const int r0 = (int)(m0 - (long long)b0 * a.Vout);
        float* Ab = A_s + buf * BMW * KROW;
        float* Gb = G_s + buf * BMW * NT;
        if (VEC) {
            constexpr int CPR = K / 4;                       // 16B cells per row
            for (int c = tid; c < BMW * CPR; c += THREADS) {
                const int mm = c / CPR, k4 = c - mm * CPR;
                float* dst = Ab + mm * KROW + 4 * k4;
                if (m0 + mm < m_end) {
                    int r = r0 + mm, b = b0;
                    while (r >= a.Vout) { r -= a.Vout; ++b; }
                    const int s = (4 * k4) / KS, co = 4 * k4 - s * KS;
                    const int src = __ldg(a.idx + r * S_ + s);
                    cp_async16(dst, a.in + ((size_t)b * a.in_rows + src) * KS + co);
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            for (int c = tid; c < BMW * K; c += THREADS) {
                const int mm = c / K, k = c - mm * K;
                float* dst = Ab + mm * KROW + k;
                if (m0 + mm < m_end) {
                    int r = r0 + mm, b = b0;
                    while (r >= a.Vout) { r -= a.Vout; ++b; }
                    const int s = k / KS, co = k - s * KS;
                    const int src = __ldg(a.idx + r * S_ + s);
                    cp_async4(dst, a.in + ((size_t)b * a.in_rows + src) * KS + co);
                } else {
                    *dst = 0.f;
                }
            }
        }
        if ((a.n_real & 3) == 0 && a.n_real == NT) {
            constexpr int Q = NT / 4;
            for (int c = tid; c < BMW * Q; c += THREADS) {
                const int mm = c / Q, q = c - mm * Q;
                float* dst = Gb + mm * NT + 4 * q;
                if (m0 + mm < m_end) cp_async16(dst, a.g + (size_t)(m0 + mm) * NT + 4 * q);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int c = tid; c < BMW * a.n_real; c += THREADS) {
                const int mm = c / a.n_real, n = c - mm * a.n_real;
                float* dst = Gb + mm * NT + n;
                if (m0 + mm < m_end) cp_async4(dst, a.g + (size_t)(m0 + mm) * a.n_real + n);
                else *dst = 0.f;
            }
        }
    };

    float acc[TK][TN];
    float accb[TN];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) accb[j] = 0.f;

    const long long nstage = (m_end > m_begin) ? (m_end - m_begin + BMW - 1) / BMW : 0;
    if (nstage > 0) { fill(m_begin, 0); cp_async_commit(); }
    for (long long st = 0; st < nstage; ++st) {
        const int buf = (int)(st & 1);
        if (st + 1 < nstage) {
            fill(m_begin + (st + 1) * BMW, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* Ab = A_s + buf * BMW * KROW + kg * (TK == 2 ? 2 : 4);
        const float* Gb = G_s + buf * BMW * NT + ng * TN;
#pragma unroll 4
        for (int mm = 0; mm < BMW; ++mm) {
            float av[TK], gv[TN];
            if (TK == 2) {
                const float2 t = *reinterpret_cast<const float2*>(Ab + mm * KROW);
                av[0] = t.x; av[1] = t.y;
            } else {
                // thread's k-set = { i*NKG*4 + kg*4 + (0..3) }: lanes stride 16 B -> conflict-free LDS.128
#pragma unroll
                for (int i = 0; i < TK / 4; ++i) {
                    const float4 t = *reinterpret_cast<const float4*>(Ab + mm * KROW + i * NKG * 4);
                    av[4 * i] = t.x; av[4 * i + 1] = t.y; av[4 * i + 2] = t.z; av[4 * i + 3] = t.w;
                }
            }
#pragma unroll
            for (int j = 0; j < TN / 4; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(Gb + mm * NT + 4 * j);
                gv[4 * j] = t.x; gv[4 * j + 1] = t.y; gv[4 * j + 2] = t.z; gv[4 * j + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TK; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
            if (kg == 0) {
#pragma unroll
                for (int j = 0; j < TN; ++j) accb[j] += gv[j];
            }
        }
        __syncthreads();
    }

    float* P = a.part + (size_t)blockIdx.x * a.n_real * K;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int n = ng * TN + j;
        if (n < a.n_real) {
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                const int k = (TK == 2) ? kg * 2 + i : (i >> 2) * NKG * 4 + kg * 4 + (i & 3);
                if (k < K) P[(size_t)n * K + k] = acc[i][j];
            }
            if (kg == 0) a.part_b[(size_t)blockIdx.x * a.n_real + n] = accb[j];
        }
    }
    (void)KP;
}

// Generic weight-gradient fallback: one thread per (n,k) element and row split.
__global__ void bw_generic_kernel(const BwArgs a, int KS, int S) {
    const int K = KS * S;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;     // element of [n_real, K+1]; k == K -> bias
    if (e >= a.n_real * (K + 1)) return;
    const int n = e / (K + 1), k = e - n * (K + 1);
    const long long m_begin = (long long)blockIdx.x * a.rows_per_cta;
    const long long m_end = min(a.M, m_begin + a.rows_per_cta);
    float acc = 0.f;
    if (k < K) {
        const int s = k / KS, c = k - s * KS;
        for (long long m = m_begin; m < m_end; ++m) {
            const int b = (int)(m / a.Vout);
            const int r = (int)(m - (long long)b * a.Vout);
            const int src = a.idx[r * S + s];
            acc = fmaf(__ldg(a.g + (size_t)m * a.n_real + n),
                       __ldg(a.in + ((size_t)b * a.in_rows + src) * KS + c), acc);
        }
        a.part[(size_t)blockIdx.x * a.n_real * K + (size_t)n * K + k] = acc;
    } else {
        for (long long m = m_begin; m < m_end; ++m) acc += __ldg(a.g + (size_t)m * a.n_real + n);
        a.part_b[(size_t)blockIdx.x * a.n_real + n] = acc;
    }
}

// out[i] = sum_{c < nsplit} part[c*len + i]   (fixed order)
__global__ void split_reduce_kernel(const float* __restrict__ part, float* __restrict__ out,
                                    int nsplit, long long len) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    float acc = 0.f;
    for (int c = 0; c < nsplit; ++c) acc += part[(size_t)c * len + i];
    out[i] = acc;
}

// The same for a weight-gradient / bias-gradient pair in ONE launch, kSplitGroups partial groups per element:
//   dW[i] = sum_c part[c*len + i] (i < len),   db[j] = sum_c part_b[c*nb + j] (j < nb)
// A block holds 64 elements x 16 groups; group g adds the partials c = g, g+16, g+32, ... in order, the group sums
// are combined by a fixed pairwise tree: the result is deterministic.  (One thread per element left 36 blocks on
// 148 SMs for a 32 x 288 weight: 10 us per layer; four groups still ran 37 dependent-latency loads per thread,
// 12.6 us per launch and seven launches a step -- 2 % of the 8-GPU step, profiles/r02_dp_floor.md.)
constexpr int kSplitGroups = 16;
__global__ void __launch_bounds__(64 * kSplitGroups)
split_reduce2_kernel(const float* __restrict__ part, const float* __restrict__ part_b,
                     float* __restrict__ dW, float* __restrict__ db, int nsplit,
                     long long len, int nb) {
    __shared__ float sm[kSplitGroups][64];
    const int e = threadIdx.x & 63, g = threadIdx.x >> 6;
    const long long i = (long long)blockIdx.x * 64 + e;
    const long long total = len + (db ? nb : 0);
    float acc = 0.f;
    if (i < total) {
        const float* src = i < len ? part + i : part_b + (i - len);
        const long long stride = i < len ? len : nb;
        for (int c = g; c < nsplit; c += kSplitGroups) acc += src[(size_t)c * stride];
    }
    sm[g][e] = acc;
    __syncthreads();
    if (g == 0 && i < total) {
        float t[kSplitGroups];
#pragma unroll
        for (int k = 0; k < kSplitGroups; ++k) t[k] = sm[k][e];
#pragma unroll
        for (int w = 1; w < kSplitGroups; w *= 2)
#pragma unroll
            for (int k = 0; k < kSplitGroups; k += 2 * w) t[k] = t[k] + t[k + w];
        if (i < len) dW[i] = t[0]; else db[i - len] = t[0];
    }
}

// Wt[c, s*Cout + o] = W[o, s*Cin + c] : the weight of the "transposed" spiral
// convolution used by the backward-to-input pass.
__global__ void weight_transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt,
                                        int Cout, int Cin, int S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = Cout * Cin * S;
    if (t >= total) return;
    const int c = t / (S * Cout);
    const int rem = t - c * S * Cout;
    const int s = rem / Cout, o = rem - s * Cout;
    Wt[t] = W[(size_t)o * S * Cin + s * Cin + c];
}

}  // namespace sdvae
