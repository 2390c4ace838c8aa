// Tensor-core variants of the spiral-convolution contractions for the wide layers
// (K = S*C_in >= 288): error-compensated 3xTF32 on mma.sync.m16n8k8.
//
// Every fp32 operand x is split in registers into  hi = x with the low 13 mantissa
// bits cleared (exactly representable in TF32)  and  lo = x - hi (exact in fp32);
// the product is accumulated in fp32 as  lo*hi' + hi*lo' + hi*hi'  (the lo*lo' term,
// <= 2^-22 relative, is dropped).  That keeps the fp32-level parity bar of the FMA
// path (measured ~1e-6 normwise) while the contraction runs on the tensor pipe.
//
// Tiling/fill are those of gc_tile_kernel (same A-tile / weight-slab staging, same
// uniform / ragged fill, same epilogue); only the inner product differs.  The k index
// inside a 32-wide chunk is permuted (thread t of a quad owns physical columns
// 8t..8t+7) so that fragments are fetched with conflict-free LDS.128.
#pragma once
#include "common.cuh"
#include "spiral_conv.cuh"

namespace sdvae {

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------
// Forward / backward-to-input / dense.  Warp tile = 64 rows x (8*NIW) columns;
// CTA = WARPS_M x WARPS_N warps -> BM = 64*WARPS_M rows, NT = 8*NIW*WARPS_N columns.
// ---------------------------------------------------------------------------
template <int KS, int NIW, int WARPS_M, int WARPS_N>
struct GmCfg {
    static constexpr int MI = 4;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int BM = 64 * WARPS_M;
    static constexpr int NT = 8 * NIW * WARPS_N;
    static constexpr int P = 36;
    static constexpr int CPS = KS / 32;
    static constexpr int CP = NT + 8;
    static size_t smem_bytes(int S, bool ragged) {
        size_t tiles = (size_t)2 * (BM + NT) * P * sizeof(float);
        size_t book = ragged ? (size_t)BM * 2 * sizeof(int) : (size_t)BM * S * sizeof(int);
        size_t epi = (size_t)BM * CP * sizeof(float);
        size_t main_part = tiles + book;
        return main_part > epi ? main_part : epi;
    }
};

template <int KS, int NIW, int WARPS_M, int WARPS_N, bool RAGGED>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32)
gc_mma_kernel(const GcArgs a) {
    using Cfg = GmCfg<KS, NIW, WARPS_M, WARPS_N>;
    constexpr int THREADS = Cfg::THREADS, BM = Cfg::BM, NT = Cfg::NT, P = Cfg::P;
    constexpr int CPS = Cfg::CPS, CP = Cfg::CP, MI = Cfg::MI;
    static_assert(KS % 32 == 0, "mma path needs 32-wide chunks");
    const int EPI = a.epi;

    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;
    float* W_s = A_s + 2 * BM * P;
    int* I_s = reinterpret_cast<int*>(W_s + 2 * NT * P);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wrow = (warp / WARPS_N) * 64, wcol = (warp % WARPS_N) * (8 * NIW);
    const long long base = (long long)blockIdx.x * BM;
    const int S = a.S;
    const int nblk = blockIdx.y * NT;
    const float* Wg = a.W + (size_t)nblk * a.ldw;

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

    float acc[MI][NIW][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NIW; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

    auto compute = [&](int buf) {
        const float* Aw = A_s + (buf * BM + wrow + g) * P + 8 * t;
        const float* Ww = W_s + (buf * NT + wcol + g) * P + 8 * t;
        uint32_t bh[NIW][8], bl[NIW][8];
#pragma unroll
        for (int ni = 0; ni < NIW; ++ni) {
            const float4 v0 = *reinterpret_cast<const float4*>(Ww + ni * 8 * P);
            const float4 v1 = *reinterpret_cast<const float4*>(Ww + ni * 8 * P + 4);
            const float w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) split_tf32(w[i], bh[ni][i], bl[ni][i]);
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            const float4 p0 = *reinterpret_cast<const float4*>(Aw + (mi * 16) * P);
            const float4 p1 = *reinterpret_cast<const float4*>(Aw + (mi * 16) * P + 4);
            const float4 q0 = *reinterpret_cast<const float4*>(Aw + (mi * 16 + 8) * P);
            const float4 q1 = *reinterpret_cast<const float4*>(Aw + (mi * 16 + 8) * P + 4);
            const float r0[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            const float r1[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t ah[4], al[4];
                split_tf32(r0[2 * j], ah[0], al[0]);        // (row g,   k = t)
                split_tf32(r1[2 * j], ah[1], al[1]);        // (row g+8, k = t)
                split_tf32(r0[2 * j + 1], ah[2], al[2]);    // (row g,   k = t+4)
                split_tf32(r1[2 * j + 1], ah[3], al[3]);    // (row g+8, k = t+4)
#pragma unroll
                for (int ni = 0; ni < NIW; ++ni) {
                    mma_tf32(acc[mi][ni], al, bh[ni][2 * j], bh[ni][2 * j + 1]);
                    mma_tf32(acc[mi][ni], ah, bl[ni][2 * j], bl[ni][2 * j + 1]);
                    mma_tf32(acc[mi][ni], ah, bh[ni][2 * j], bh[ni][2 * j + 1]);
                }
            }
        }
    };

    constexpr int NCELL = BM * 8 / THREADS;
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
    auto a_fill_ragged = [&](int ch, int buf) {
        // cells with exactly one source row go through cp.async; empty cells are zeroed;
        // only cells with >= 2 source rows (about a fifth) are summed through registers
        const int s = ch / CPS, h = ch - s * CPS;
#pragma unroll
        for (int i = 0; i < NCELL; ++i) {
            const int c = tid + i * THREADS;
            const int lr = c >> 3, q = c & 7;
            float* dst = A_s + (buf * BM + lr) * P + 4 * q;
            const int r = I_s[2 * lr + 1];
            int e0 = 0, e1 = 0;
            if (r >= 0) {
                const int cell = r * S + s;
                e0 = __ldg(a.cell_ptr + cell);
                e1 = __ldg(a.cell_ptr + cell + 1);
            }
            const float* rowbase = a.in + (size_t)I_s[2 * lr] * KS + h * 32 + 4 * q;
            if (e1 - e0 == 1) {
                cp_async16(dst, rowbase + (size_t)__ldg(a.cell_src + e0) * KS);
            } else {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int e = e0; e < e1; ++e) {
                    const float4 u = ldg4(rowbase + (size_t)__ldg(a.cell_src + e) * KS);
                    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                }
                *reinterpret_cast<float4*>(dst) = v;
            }
        }
    };
    auto a_fill = [&](int ch, int buf) {
        if constexpr (RAGGED) a_fill_ragged(ch, buf);
        else a_fill_async(ch, buf);
    };

    a_fill(0, 0);
    w_fill(0, 0);
    cp_async_commit();
    for (int ch = 0; ch < NCH; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < NCH) {
            a_fill(ch + 1, buf ^ 1);
            w_fill(ch + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        compute(buf);
        __syncthreads();
    }

    // ---- epilogue -------------------------------------------------------------------
    float* C_s = smem;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NIW; ++ni) {
            const int col = wcol + ni * 8 + 2 * t;
            float b0 = 0.f, b1 = 0.f;
            if ((EPI == EPI_BIAS || EPI == EPI_BIAS_ELU) && a.bias) {
                if (nblk + col < a.n_real) b0 = __ldg(a.bias + nblk + col);
                if (nblk + col + 1 < a.n_real) b1 = __ldg(a.bias + nblk + col + 1);
            }
            float v0 = acc[mi][ni][0] + b0, v1 = acc[mi][ni][1] + b1;
            float v2 = acc[mi][ni][2] + b0, v3 = acc[mi][ni][3] + b1;
            if (EPI == EPI_BIAS_ELU) { v0 = elu_f(v0); v1 = elu_f(v1); v2 = elu_f(v2); v3 = elu_f(v3); }
            const int row = wrow + mi * 16 + g;
            *reinterpret_cast<float2*>(C_s + row * CP + col) = make_float2(v0, v1);
            *reinterpret_cast<float2*>(C_s + (row + 8) * CP + col) = make_float2(v2, v3);
        }
    __syncthreads();
    const int ncols = min(NT, a.n_real - nblk);
    if (ncols == NT && (a.ldo & 3) == 0) {
        constexpr int Q = NT / 4;
        for (int e = tid; e < BM * Q; e += THREADS) {
            const int row = e / Q, c4 = e - row * Q;
            const long long m = base + row;
            if (m < a.M) {
                float4 v = *reinterpret_cast<const float4*>(C_s + row * CP + 4 * c4);
                const size_t off = (size_t)m * a.ldo + nblk + 4 * c4;
                if (EPI == EPI_GATE) {
                    const float4 gt = ldg4(a.gate + off);
                    v.x *= elu_grad_from_out(gt.x); v.y *= elu_grad_from_out(gt.y);
                    v.z *= elu_grad_from_out(gt.z); v.w *= elu_grad_from_out(gt.w);
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
// Weight gradient on the tensor pipe:  dW^T[k, n] = sum_m A[m, k] * g[m, n].
// One warp per 32-wide k block (K/32 warps per CTA), all NT columns; the reduction
// dimension m streams through shared memory 32 rows at a time.
// ---------------------------------------------------------------------------
template <int KS, int S_, int NT>
struct BmCfg {
    static constexpr int K = KS * S_;
    static constexpr int NWARPS = K / 32;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int BMW = 32;
    static constexpr int KROW = K + 8;      // == 8 (mod 32): conflict-free LDS.128 across rows
    static constexpr int GP = NT + 8;
    static constexpr int NI = NT / 8;
    static constexpr size_t SMEM = (size_t)2 * BMW * (KROW + GP) * sizeof(float);
};

template <int KS, int S_, int NT>
__global__ void __launch_bounds__(BmCfg<KS, S_, NT>::THREADS)
bw_mma_kernel(const BwArgs a) {
    using Cfg = BmCfg<KS, S_, NT>;
    constexpr int K = Cfg::K, THREADS = Cfg::THREADS, BMW = Cfg::BMW, KROW = Cfg::KROW;
    constexpr int GP = Cfg::GP, NI = Cfg::NI;
    static_assert(K % 32 == 0 && KS % 4 == 0 && (NT == 32 || NT == 64), "bw_mma shape");

    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;                         // [2][BMW][KROW]
    float* G_s = smem + 2 * BMW * KROW;        // [2][BMW][GP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int kb = warp * 32;
    const long long m_begin = (long long)blockIdx.x * a.rows_per_cta;
    const long long m_end = min(a.M, m_begin + a.rows_per_cta);

    auto fill = [&](long long m0, int buf) {
        const int b0 = (int)(m0 / a.Vout);
        const int r0 = (int)(m0 - (long long)b0 * a.Vout);
        float* Ab = A_s + buf * BMW * KROW;
        float* Gb = G_s + buf * BMW * GP;
        constexpr int CPR = K / 4;
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
        constexpr int Q = NT / 4;
        for (int c = tid; c < BMW * Q; c += THREADS) {
            const int mm = c / Q, q = c - mm * Q;
            float* dst = Gb + mm * GP + 4 * q;
            if (m0 + mm < m_end) cp_async16(dst, a.g + (size_t)(m0 + mm) * NT + 4 * q);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    float acc[2][NI][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
    float bsum = 0.f;

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
        const float* Ab = A_s + buf * BMW * KROW + kb + 4 * g;
        const float* Gb = G_s + buf * BMW * GP + (NI == 4 ? 4 : 8) * g;
#pragma unroll
        for (int ms = 0; ms < BMW / 8; ++ms) {
            const int m0 = ms * 8;
            // A fragments: physical k = kb + 4g + {0,1,2,3}; tile i: row g <-> 2i, row g+8 <-> 2i+1
            const float4 x0 = *reinterpret_cast<const float4*>(Ab + (m0 + t) * KROW);
            const float4 x1 = *reinterpret_cast<const float4*>(Ab + (m0 + t + 4) * KROW);
            const float xa[4] = {x0.x, x0.y, x0.z, x0.w};
            const float xb[4] = {x1.x, x1.y, x1.z, x1.w};
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                split_tf32(xa[2 * i], ah[i][0], al[i][0]);        // (row g,   m = t)
                split_tf32(xa[2 * i + 1], ah[i][1], al[i][1]);    // (row g+8, m = t)
                split_tf32(xb[2 * i], ah[i][2], al[i][2]);        // (row g,   m = t+4)
                split_tf32(xb[2 * i + 1], ah[i][3], al[i][3]);    // (row g+8, m = t+4)
            }
            // B fragments: fragment column g of n-tile ni <-> physical n = NI*g + ni
            float gv0[NI], gv1[NI];
#pragma unroll
            for (int q = 0; q < NI / 4; ++q) {
                const float4 y0 = *reinterpret_cast<const float4*>(Gb + (m0 + t) * GP + 4 * q);
                const float4 y1 = *reinterpret_cast<const float4*>(Gb + (m0 + t + 4) * GP + 4 * q);
                gv0[4 * q] = y0.x; gv0[4 * q + 1] = y0.y; gv0[4 * q + 2] = y0.z; gv0[4 * q + 3] = y0.w;
                gv1[4 * q] = y1.x; gv1[4 * q + 1] = y1.y; gv1[4 * q + 2] = y1.z; gv1[4 * q + 3] = y1.w;
            }
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(gv0[ni], bh0, bl0);
                split_tf32(gv1[ni], bh1, bl1);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    mma_tf32(acc[i][ni], al[i], bh0, bh1);
                    mma_tf32(acc[i][ni], ah[i], bl0, bl1);
                    mma_tf32(acc[i][ni], ah[i], bh0, bh1);
                }
            }
        }
        if (tid < NT) {
            const float* Gc = G_s + buf * BMW * GP + tid;
#pragma unroll 8
            for (int mm = 0; mm < BMW; ++mm) bsum += Gc[mm * GP];
        }
        __syncthreads();
    }

    // acc[i][ni][c]: row g (c<2) / g+8 (c>=2) of k-tile i -> k = kb + 4g + 2i + (c>>1);
    //                column 2t + (c&1) of n-tile ni       -> n = NI*(2t + (c&1)) + ni
    float* P = a.part + (size_t)blockIdx.x * NT * K;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = kb + 4 * g + 2 * i + (c >> 1);
                const int n = NI * (2 * t + (c & 1)) + ni;
                P[(size_t)n * K + k] = acc[i][ni][c];
            }
    if (tid < NT) a.part_b[(size_t)blockIdx.x * NT + tid] = bsum;
}

}  // namespace sdvae
