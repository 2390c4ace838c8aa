// Loss reductions of the training step (reference model_manager.py:281-312):
//   reconstruction MSE (:332-334), Laplacian regulariser (:343-349, utils.py:153-165),
//   KL divergence (:351-354), latent consistency (:360-393).
// Every reduction is a fixed tree (warp shuffles -> per-block partial -> one
// finishing block in double), so repeated runs give identical bits.
#pragma once
#include "common.cuh"

namespace sdvae {

constexpr int kLossThreads = 256;

// ---- MSE + Laplacian, forward ------------------------------------------------
// One thread per (b, v):  q = sum_j lval[v,j] * recon[b, lcol[v,j], :]  (ELL, -1 pad),
// writes qn = q / |q| (0 where |q| == 0) for the backward pass and per-block
// partial sums of  (recon-x)^2  and  |q|.
__global__ void __launch_bounds__(kLossThreads)
mse_lap_fwd_kernel(const float* __restrict__ recon, const float* __restrict__ x,
                   const int* __restrict__ lcol, const float* __restrict__ lval, int lw,
                   float* __restrict__ qn, float* __restrict__ partial /* [grid][2] */,
                   long long BV, int V) {
    __shared__ float scratch[kLossThreads / 32];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float se = 0.f, nq = 0.f;
    if (t < BV) {
        const int v = (int)(t % V);
        const long long b = t / V;
        const float* rb = recon + (size_t)b * V * 3;
        const float r0 = rb[v * 3], r1 = rb[v * 3 + 1], r2 = rb[v * 3 + 2];
        const float* xr = x + (size_t)t * 3;
        const float d0 = r0 - xr[0], d1 = r1 - xr[1], d2 = r2 - xr[2];
        se = d0 * d0 + d1 * d1 + d2 * d2;
        if (lcol) {
            float q0 = 0.f, q1 = 0.f, q2 = 0.f;
            for (int j = 0; j < lw; ++j) {
                const int c = __ldg(lcol + v * lw + j);
                if (c < 0) continue;
                const float w = __ldg(lval + v * lw + j);
                q0 = fmaf(w, rb[c * 3], q0);
                q1 = fmaf(w, rb[c * 3 + 1], q1);
                q2 = fmaf(w, rb[c * 3 + 2], q2);
            }
            nq = sqrtf(q0 * q0 + q1 * q1 + q2 * q2);
            const float inv = nq > 0.f ? 1.f / nq : 0.f;
            float* o = qn + (size_t)t * 3;
            o[0] = q0 * inv; o[1] = q1 * inv; o[2] = q2 * inv;
        }
    }
    const float s0 = block_sum<kLossThreads>(se, scratch);
    const float s1 = block_sum<kLossThreads>(nq, scratch);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s0; partial[2 * blockIdx.x + 1] = s1; }
}

// ---- MSE + Laplacian, backward -------------------------------------------------
//   drecon[b,u,:] = c_mse * 2 (recon - x)  +  c_lap * sum_{e in T(u)} tval[e] * qn[b, trow[e], :]
// T = transposed Laplacian in CSR (the random-walk Laplacian is not symmetric).
// c_mse = g_mse / (B*V*3), c_lap = g_lap / (B*V); g_* are upstream gradients
// (the loss weights), optionally multiplied by device scalars dscale[0], dscale[1].
__global__ void mse_lap_bwd_kernel(const float* __restrict__ recon, const float* __restrict__ x,
                                   const float* __restrict__ qn, const int* __restrict__ tptr,
                                   const int* __restrict__ trow, const float* __restrict__ tval,
                                   float* __restrict__ drecon, long long BV, int V,
                                   float c_mse, float c_lap, const float* __restrict__ dscale) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= BV) return;
    if (dscale) { c_mse *= dscale[0]; c_lap *= dscale[1]; }
    const int u = (int)(t % V);
    const long long b = t / V;
    const float* rr = recon + (size_t)t * 3;
    const float* xr = x + (size_t)t * 3;
    float g0 = 2.f * c_mse * (rr[0] - xr[0]);
    float g1 = 2.f * c_mse * (rr[1] - xr[1]);
    float g2 = 2.f * c_mse * (rr[2] - xr[2]);
    if (tptr) {
        const float* qb = qn + (size_t)b * V * 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const int e0 = __ldg(tptr + u), e1 = __ldg(tptr + u + 1);
        for (int e = e0; e < e1; ++e) {
            const int r = __ldg(trow + e);
            const float w = __ldg(tval + e);
            a0 = fmaf(w, qb[r * 3], a0);
            a1 = fmaf(w, qb[r * 3 + 1], a1);
            a2 = fmaf(w, qb[r * 3 + 2], a2);
        }
        g0 = fmaf(c_lap, a0, g0); g1 = fmaf(c_lap, a1, g1); g2 = fmaf(c_lap, a2, g2);
    }
    float* o = drecon + (size_t)t * 3;
    o[0] = g0; o[1] = g1; o[2] = g2;
}

// ---- KL ---------------------------------------------------------------------------
//   L = (1/B) sum_b -1/2 sum_d (1 + lv - mu^2 - e^lv);  dmu = mu/B;  dlv = (e^lv - 1)/(2B)
__global__ void __launch_bounds__(kLossThreads)
kl_fwd_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                  float* __restrict__ dmu, float* __restrict__ dlv,
                  float* __restrict__ partial /* [grid] */, long long n, float inv_b) {
    __shared__ float scratch[kLossThreads / 32];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float term = 0.f;
    if (i < n) {
        const float m = mu[i], l = lv[i], e = expf(l);
        term = -0.5f * (1.f + l - m * m - e);
        dmu[i] = m * inv_b;
        dlv[i] = 0.5f * (e - 1.f) * inv_b;
    }
    const float s = block_sum<kLossThreads>(term, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// ---- latent consistency --------------------------------------------------------------
// z is [bs*bs, D] on the swap grid (row i*bs+j = base mesh i, feature donor j).
// zf = columns [r0,r1), ze = the rest.  For pair a<b and index t:
//   h2 = max(0, |ze[t,b]-ze[t,a]|^2 - |ze[b,t]-ze[a,t]|^2 + eta2)
//   h1 = max(0, |zf[b,t]-zf[a,t]|^2 - |zf[t,b]-zf[t,a]|^2 + eta1)
// Pass 1: one thread per (pair, t) -> hinge values, activity flags, block partials.
__device__ __forceinline__ void lc_dist(const float* __restrict__ p, const float* __restrict__ q,
                                        int D, int r0, int r1, float& df, float& de) {
    float f = 0.f, e = 0.f;
    for (int d = 0; d < D; ++d) {
        const float t = p[d] - q[d];
        if (d >= r0 && d < r1) f = fmaf(t, t, f); else e = fmaf(t, t, e);
    }
    df = f; de = e;
}

__global__ void __launch_bounds__(kLossThreads)
lc_hinge_kernel(const float* __restrict__ z, int bs, int D, int r0, int r1, float eta1, float eta2,
                unsigned char* __restrict__ act /* [npairs*bs][2] */, float* __restrict__ partial) {
    __shared__ float scratch[kLossThreads / 32];
    const int npairs = bs * (bs - 1) / 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float h = 0.f;
    if (i < (long long)npairs * bs) {
        const int pair = (int)(i / bs), t = (int)(i % bs);
        // decode pair -> (a, b), a < b, in row-major upper-triangular order
        int a = 0, rem = pair;
        while (rem >= bs - 1 - a) { rem -= bs - 1 - a; ++a; }
        const int b = a + 1 + rem;
        float f_col, e_col, f_row, e_row;
        lc_dist(z + (size_t)(b * bs + t) * D, z + (size_t)(a * bs + t) * D, D, r0, r1, f_col, e_col);
        lc_dist(z + (size_t)(t * bs + b) * D, z + (size_t)(t * bs + a) * D, D, r0, r1, f_row, e_row);
        const float h2 = e_row - e_col + eta2;     // lr - dr + eta2
        const float h1 = f_col - f_row + eta1;     // lg - dg + eta1
        act[2 * i] = h1 > 0.f;
        act[2 * i + 1] = h2 > 0.f;
        h = fmaxf(h1, 0.f) + fmaxf(h2, 0.f);
    }
    const float s = block_sum<kLossThreads>(h, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Pass 2: one thread per (grid row (i,j), latent d) gathers every active hinge that
// touches z[i*bs+j, d] -- an owner-computes gather, no atomics.
__device__ __forceinline__ int lc_pair_index(int a, int b, int bs) {   // a < b
    return a * (bs - 1) - a * (a - 1) / 2 + (b - a - 1);
}

__global__ void lc_grad_kernel(const float* __restrict__ z, const unsigned char* __restrict__ act,
                               int bs, int D, int r0, int r1, float scale /* 2/(bs^3-bs^2) */,
                               float* __restrict__ dz) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)bs * bs * D) return;
    const int d = (int)(idx % D);
    const int row = (int)(idx / D);
    const int i = row / bs, j = row % bs;
    const bool feat = d >= r0 && d < r1;
    const int which = feat ? 0 : 1;
    // feature dims: +col-direction (lg), -row-direction (dg); other dims: -col (dr), +row (lr)
    const float s_col = feat ? 1.f : -1.f;
    const float zc = z[(size_t)row * D + d];
    float g = 0.f;
    for (int o = 0; o < bs; ++o) {
        if (o != i) {           // column direction: rows (i,j) and (o,j), pair {i,o}, t = j
            const int p = i < o ? lc_pair_index(i, o, bs) : lc_pair_index(o, i, bs);
            if (act[2 * ((size_t)p * bs + j) + which])
                g += s_col * (zc - z[(size_t)(o * bs + j) * D + d]);
        }
        if (o != j) {           // row direction: rows (i,j) and (i,o), pair {j,o}, t = i
            const int p = j < o ? lc_pair_index(j, o, bs) : lc_pair_index(o, j, bs);
            if (act[2 * ((size_t)p * bs + i) + which])
                g -= s_col * (zc - z[(size_t)(i * bs + o) * D + d]);
        }
    }
    dz[idx] = scale * g;
}

// ---- finishing reduction ------------------------------------------------------------------
// L1 reconstruction loss (reference model_manager.py:328-330, torch.nn.L1Loss(reduction='mean')): per-block partial
// sums of |a - b| (warp shuffle -> block), fixed order; l1_bwd_kernel: d/da = sign(a - b) * scale.
__global__ void __launch_bounds__(kLossThreads)
l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ partial, long long n) {
    __shared__ float sh[kLossThreads / 32];
    const long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x;
    float v = i < n ? fabsf(a[i] - b[i]) : 0.f;
    v = block_sum<kLossThreads>(v, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}

__global__ void l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ da,
                              long long n, float scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = a[i] - b[i];
    da[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
}

// Sums `count` floats with stride `stride` starting at part[offset] in double, in one block.
// out[slot] = scale * sum.
__global__ void __launch_bounds__(kLossThreads)
finish_sum_kernel(const float* __restrict__ part, long long count, int stride, int offset,
                  float scale, float* __restrict__ out, int slot) {
    __shared__ double sh[kLossThreads / 32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < count; i += kLossThreads) acc += (double)part[i * stride + offset];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double r = threadIdx.x < kLossThreads / 32 ? sh[threadIdx.x] : 0.0;
        r = warp_sum_d(r);
        if (threadIdx.x == 0) out[slot] = (float)(r * (double)scale);
    }
}

// losses[6] = losses[0] + w_kl*losses[1] + w_lc*losses[2] + w_lap*losses[3] (+ w_cls*losses[4])
// (model_manager.py:308-312; slot order = ModelManager.loss_keys, :150-154)
__global__ void total_loss_kernel(float* losses, float w_kl, float w_lc, float w_lap, float w_cls) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        losses[6] = losses[0] + w_kl * losses[1] + w_lc * losses[2] + w_lap * losses[3] + w_cls * losses[4];
}

// out = a + sa * b  (+ sc * c)   -- merges the latent-space gradients
__global__ void axpy3_kernel(const float* __restrict__ a, const float* __restrict__ b, float sb,
                             const float* __restrict__ c, float sc, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = a ? a[i] : 0.f;
    if (b) v = fmaf(sb, b[i], v);
    if (c) v = fmaf(sc, c[i], v);
    out[i] = v;
}

}  // namespace sdvae
