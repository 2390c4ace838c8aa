// Pool (reference model.py:50-55), the feature swap (swap_batch_transform.py:13-52),
// re-parameterisation (model.py:184-188), ELU-backward gate and Adam
// (model_manager.py:69-72, 316).  All memory-bound; one thread per 16 bytes.
#pragma once
#include "common.cuh"

namespace sdvae {

// ---- Pool forward: fixed-width ELL rows ------------------------------------
//   out[b,r,:] = sum_{j<Wd, col[r,j]>=0} val[r,j] * x[b, col[r,j], :]
// entries in storage order; product rounded, then added (separate mul/add, like
// the reference's `index_select * value` followed by scatter_add onto zeros).
template <bool VEC>
__global__ void pool_ell_fwd_kernel(const float* __restrict__ x, const int* __restrict__ col,
                                    const float* __restrict__ val, float* __restrict__ out,
                                    long long total /* B*Vout*C/(VEC?4:1) */, int Vin, int Vout,
                                    int Wd, int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int CQ = VEC ? C / 4 : C;
    const int cq = (int)(t % CQ);
    const long long br = t / CQ;
    const int r = (int)(br % Vout);
    const long long b = br / Vout;
    const float* xb = x + (size_t)b * Vin * C;
    if (VEC) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < Wd; ++j) {
            const int cj = __ldg(col + r * Wd + j);
            if (cj < 0) continue;
            const float w = __ldg(val + r * Wd + j);
            const float4 v = ldg4(xb + (size_t)cj * C + 4 * cq);
            acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, w));
            acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, w));
            acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, w));
            acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, w));
        }
        *reinterpret_cast<float4*>(out + (size_t)br * C + 4 * cq) = acc;
    } else {
        float acc = 0.f;
        for (int j = 0; j < Wd; ++j) {
            const int cj = __ldg(col + r * Wd + j);
            if (cj < 0) continue;
            acc = __fadd_rn(acc, __fmul_rn(__ldg(xb + (size_t)cj * C + cq), __ldg(val + r * Wd + j)));
        }
        out[(size_t)br * C + cq] = acc;
    }
}

// ---- Pool backward / generic CSR row-sum -----------------------------------
//   dx[b,k,:] = gate( sum_{e in [ptr[k],ptr[k+1])} val[e] * dy[b, src[e], :] )
// One owner thread per output element, entries visited in stored order:
// deterministic, atomic-free.  val == nullptr means all-ones (used to scatter
// the encoder's per-slot input gradients).  gate != nullptr multiplies by
// elu'(gate) (the activation that produced the pooled tensor).
template <bool VEC>
__global__ void csr_rowsum_kernel(const float* __restrict__ dy, const int* __restrict__ ptr,
                                  const int* __restrict__ src, const float* __restrict__ val,
                                  const float* __restrict__ gate, float* __restrict__ dx,
                                  long long total, int Vsrc, int Vdst, int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int CQ = VEC ? C / 4 : C;
    const int cq = (int)(t % CQ);
    const long long bk = t / CQ;
    const int k = (int)(bk % Vdst);
    const long long b = bk / Vdst;
    const float* dyb = dy + (size_t)b * Vsrc * C;
    const int e0 = __ldg(ptr + k), e1 = __ldg(ptr + k + 1);
    if (VEC) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = e0; e < e1; ++e) {
            const float w = val ? __ldg(val + e) : 1.f;
            const float4 v = ldg4(dyb + (size_t)__ldg(src + e) * C + 4 * cq);
            acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, w));
            acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, w));
            acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, w));
            acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, w));
        }
        const size_t off = (size_t)bk * C + 4 * cq;
        if (gate) {
            const float4 g = ldg4(gate + off);
            acc.x *= elu_grad_from_out(g.x); acc.y *= elu_grad_from_out(g.y);
            acc.z *= elu_grad_from_out(g.z); acc.w *= elu_grad_from_out(g.w);
        }
        *reinterpret_cast<float4*>(dx + off) = acc;
    } else {
        float acc = 0.f;
        for (int e = e0; e < e1; ++e)
            acc = __fadd_rn(acc, __fmul_rn(__ldg(dyb + (size_t)__ldg(src + e) * C + cq),
                                           val ? __ldg(val + e) : 1.f));
        const size_t off = (size_t)bk * C + cq;
        if (gate) acc *= elu_grad_from_out(__ldg(gate + off));
        dx[off] = acc;
    }
}

// ---- the same two kernels, kPoolMeshes meshes per thread ---------------------------------------
// The (col, val) / (ptr, src, val) entries of an output row are the same for every mesh of the batch: one
// thread computes its 16-byte piece for kPoolMeshes meshes, loading the entries once and keeping
// kPoolMeshes x (entries) independent gathers in flight.  (The one-mesh-per-thread kernels ran at 36 % of
// the measured HBM peak at level 0: two dependent loads per thread and nothing to overlap them with.)
// Per mesh the arithmetic and its order are unchanged (bit-exact against the reference CPU result).
constexpr int kPoolMeshes = 4;

__global__ void pool_ell_fwd_batch_kernel(const float* __restrict__ x, const int* __restrict__ col,
                                          const float* __restrict__ val, float* __restrict__ out,
                                          long long total /* ceil(B/kPoolMeshes)*Vout*C/4 */, int B,
                                          int Vin, int Vout, int Wd, int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int CQ = C / 4;
    const int cq = (int)(t % CQ);
    const long long gr = t / CQ;
    const int r = (int)(gr % Vout);
    const int b0 = (int)(gr / Vout) * kPoolMeshes;
    float4 acc[kPoolMeshes];
#pragma unroll
    for (int i = 0; i < kPoolMeshes; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < Wd; ++j) {
        const int cj = __ldg(col + r * Wd + j);
        if (cj < 0) continue;
        const float w = __ldg(val + r * Wd + j);
        float4 v[kPoolMeshes];
#pragma unroll
        for (int i = 0; i < kPoolMeshes; ++i)
            v[i] = b0 + i < B ? ldg4(x + ((size_t)(b0 + i) * Vin + cj) * C + 4 * cq) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kPoolMeshes; ++i) {
            acc[i].x = __fadd_rn(acc[i].x, __fmul_rn(v[i].x, w));
            acc[i].y = __fadd_rn(acc[i].y, __fmul_rn(v[i].y, w));
            acc[i].z = __fadd_rn(acc[i].z, __fmul_rn(v[i].z, w));
            acc[i].w = __fadd_rn(acc[i].w, __fmul_rn(v[i].w, w));
        }
    }
#pragma unroll
    for (int i = 0; i < kPoolMeshes; ++i)
        if (b0 + i < B) *reinterpret_cast<float4*>(out + ((size_t)(b0 + i) * Vout + r) * C + 4 * cq) = acc[i];
}

__global__ void csr_rowsum_batch_kernel(const float* __restrict__ dy, const int* __restrict__ ptr,
                                        const int* __restrict__ src, const float* __restrict__ val,
                                        const float* __restrict__ gate, float* __restrict__ dx,
                                        long long total, int B, int Vsrc, int Vdst, int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int CQ = C / 4;
    const int cq = (int)(t % CQ);
    const long long gk = t / CQ;
    const int k = (int)(gk % Vdst);
    const int b0 = (int)(gk / Vdst) * kPoolMeshes;
    const int e0 = __ldg(ptr + k), e1 = __ldg(ptr + k + 1);
    float4 acc[kPoolMeshes];
#pragma unroll
    for (int i = 0; i < kPoolMeshes; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
        const float w = val ? __ldg(val + e) : 1.f;
        const int sr = __ldg(src + e);
        float4 v[kPoolMeshes];
#pragma unroll
        for (int i = 0; i < kPoolMeshes; ++i)
            v[i] = b0 + i < B ? ldg4(dy + ((size_t)(b0 + i) * Vsrc + sr) * C + 4 * cq) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kPoolMeshes; ++i) {
            acc[i].x = __fadd_rn(acc[i].x, __fmul_rn(v[i].x, w));
            acc[i].y = __fadd_rn(acc[i].y, __fmul_rn(v[i].y, w));
            acc[i].z = __fadd_rn(acc[i].z, __fmul_rn(v[i].z, w));
            acc[i].w = __fadd_rn(acc[i].w, __fmul_rn(v[i].w, w));
        }
    }
#pragma unroll
    for (int i = 0; i < kPoolMeshes; ++i) {
        if (b0 + i >= B) continue;
        const size_t off = ((size_t)(b0 + i) * Vdst + k) * C + 4 * cq;
        float4 a = acc[i];
        if (gate) {
            const float4 g = ldg4(gate + off);
            a.x *= elu_grad_from_out(g.x); a.y *= elu_grad_from_out(g.y);
            a.z *= elu_grad_from_out(g.z); a.w *= elu_grad_from_out(g.w);
        }
        *reinterpret_cast<float4*>(dx + off) = a;
    }
}

// ---- elementwise -------------------------------------------------------------
// out = dy * elu'(y)
__global__ void elu_gate_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = dy[i] * elu_grad_from_out(y[i]);
}
// y = elu(x)
__global__ void elu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = elu_f(x[i]);
}

// z = mu + eps * exp(0.5 * logvar)                       (model.py:184-188)
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                   const float* __restrict__ eps, float* __restrict__ z, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = mu[i] + eps[i] * expf(0.5f * logvar[i]);
}
// dmu += dz ; dlogvar += dz * eps * 0.5 * exp(0.5*logvar)
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ logvar,
                                   const float* __restrict__ eps, float* __restrict__ dmu,
                                   float* __restrict__ dlogvar, long long n, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = dz[i];
    const float dl = g * eps[i] * 0.5f * expf(0.5f * logvar[i]);
    if (accumulate) { dmu[i] += g; dlogvar[i] += dl; }
    else { dmu[i] = g; dlogvar[i] = dl; }
}

// ---- feature swap --------------------------------------------------------------
// out[(i - i0)*bs + j, v, :] = mask[v] ? x[j, v, :] : x[i, v, :]   for i in [i0, i1)
// (swap_batch_transform.py:27-38: element i*bs+j is base mesh i with the swapped
// region taken from mesh j; the diagonal is the original).
__global__ void swap_kernel(const float* __restrict__ x, const unsigned char* __restrict__ mask,
                            float* __restrict__ out, int bs, int i0, int i1, int V, int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_mesh = (long long)V * C;
    const long long total = (long long)(i1 - i0) * bs * per_mesh;
    if (t >= total) return;
    const long long mesh = t / per_mesh;
    const long long rem = t - mesh * per_mesh;
    const int v = (int)(rem / C);
    const int i = i0 + (int)(mesh / bs), j = (int)(mesh % bs);
    const int srcm = mask[v] ? j : i;
    out[t] = __ldg(x + (size_t)srcm * per_mesh + rem);
}

// ---- Adam (torch.optim.Adam, amsgrad=False, maximize=False) ---------------------
//   g  = gscale * grad (+ wd * p)
//   m  = m + (1-b1) * (g - m)          (torch: exp_avg.lerp_(g, 1-b1))
//   v  = b2 * v + (1-b2) * g * g
//   p -= (lr / (1-b1^t)) * m / ( sqrt(v) / sqrt(1-b2^t) + eps )
// `step` lives on the device (incremented by adam_tick_kernel) so that the whole
// training step can be replayed from a CUDA graph.
__global__ void adam_tick_kernel(int* step) { if (threadIdx.x == 0 && blockIdx.x == 0) *step += 1; }

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ grad,
                            float* __restrict__ m, float* __restrict__ v, long long n,
                            const int* __restrict__ step_dev, int step_host, float lr, float b1,
                            float b2, float eps, float wd, float gscale) {
    __shared__ float s_step_size, s_sqrt_bc2;
    if (threadIdx.x == 0) {
        const int t = step_dev ? *step_dev : step_host;
        const double bc1 = 1.0 - pow((double)b1, (double)t);
        const double bc2 = 1.0 - pow((double)b2, (double)t);
        s_step_size = (float)((double)lr / bc1);
        s_sqrt_bc2 = (float)sqrt(bc2);
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = grad[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) g = fmaf(wd, pi, g);
    float mi = m[i], vi = v[i];
    mi = mi + (1.f - b1) * (g - mi);
    vi = b2 * vi + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / s_sqrt_bc2 + eps;
    p[i] = pi - s_step_size * (mi / denom);
}

}  // namespace sdvae
