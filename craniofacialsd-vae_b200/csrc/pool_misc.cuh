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

// ---- Pool forward with the tile's source rows staged in shared memory --------------------------
// The ELL gather above reads W source rows per output row through L2 (3 x 128 B per 128 B written at the
// up-sampling levels): it runs at the L2->SM fabric ceiling (~7 TB/s of gathered rows), i.e. 47 % of the
// HBM peak in algorithmic bytes.  But a tile of kPoolTile consecutive output rows touches few DISTINCT
// source rows (up-sampling: ~0.9 per output row against 3 entries; every coarse vertex feeds ~12 fine ones),
// so a host-built plan (tables.pool_stage_plan) lists them per tile; the CTA copies them once per mesh with
// cp.async into a ring of NST buffers (NST-1 meshes in flight) and the W reads per output row
// become conflict-free LDS.128 (the CQ lanes of an output row read one whole staged row).  A thread keeps its
// rows' entries and its share of the copy addresses in registers for all the meshes of its CTA.  Per mesh the
// arithmetic and its order are those of pool_ell_fwd_kernel: bit-exact against the reference CPU result.
//   tile_ptr [L+1], stage_src [tile_ptr[L]]: source rows staged by tile t, ascending
//   ent [Vout, WD] int2: .x = position of the entry's source row in its tile's staged list (-1 = padding),
//                        .y = bits of the fp32 value
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

constexpr int kPoolStageThreads = 256;
constexpr int kPoolTile = 128;            // output rows per tile (= tables.POOL_STAGE_TILE)
constexpr int pool_stages_for(int CQ) { return CQ == 8 ? 3 : 4; }      // ring depth (see launch_pool_staged)
constexpr int kPoolMaxIssue = 8;          // cp.async per thread and mesh: ucap * CQ <= 8 * 256

template <int CQ, int WD, int NST>
__global__ void __launch_bounds__(kPoolStageThreads)
pool_ell_fwd_staged_kernel(const float* __restrict__ x, const int* __restrict__ tile_ptr,
                           const int* __restrict__ stage_src, const int2* __restrict__ ent,
                           float* __restrict__ out, int B, int Vin, int Vout, int L, int ucap, int MG) {
    constexpr int RPT = kPoolTile * CQ / kPoolStageThreads;     // output pieces per thread and mesh
    constexpr int RSTEP = kPoolStageThreads / CQ;               // tile rows between a thread's pieces
    extern __shared__ float4 pool_stage[];                      // [NST][ucap * CQ]
    const int tile = (int)blockIdx.x % L, grp = (int)blockIdx.x / L;
    const int m0 = grp * MG, m1 = min(B, m0 + MG);
    if (m0 >= m1) return;
    const int u0 = __ldg(tile_ptr + tile), U = __ldg(tile_ptr + tile + 1) - u0;
    const int r0 = tile * kPoolTile;
    const int tid = threadIdx.x, q = tid % CQ, lr0 = tid / CQ;
    const size_t mesh_in = (size_t)Vin * CQ, mesh_out = (size_t)Vout * CQ;     // in float4
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* out4 = reinterpret_cast<float4*>(out);
    const int stage_f4 = ucap * CQ;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(pool_stage);

    int src_off[kPoolMaxIssue];                                 // float4 offset inside a mesh, -1 = nothing
#pragma unroll
    for (int i = 0; i < kPoolMaxIssue; ++i) {
        const int p = tid + i * kPoolStageThreads;
        src_off[i] = p < U * CQ ? __ldg(stage_src + u0 + p / CQ) * CQ + q : -1;
    }
    int loc[RPT][WD];
    float wv[RPT][WD];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int r = r0 + lr0 + k * RSTEP;
#pragma unroll
        for (int j = 0; j < WD; ++j) {
            const int2 e = r < Vout ? __ldg(ent + (size_t)r * WD + j) : make_int2(-1, 0);
            loc[k][j] = e.x < 0 ? -1 : (e.x * CQ + q) * 16;           // byte offset inside a stage buffer
            wv[k][j] = __int_as_float(e.y);
        }
    }

    auto issue = [&](int m, int buf) {
        if (m < m1) {
            const float4* xm = x4 + (size_t)m * mesh_in;
            float4* dst = pool_stage + (size_t)buf * stage_f4 + tid;
#pragma unroll
            for (int i = 0; i < kPoolMaxIssue; ++i)
                if (src_off[i] >= 0) cp_async16(dst + i * kPoolStageThreads, xm + src_off[i]);
        }
        cp_async_commit();                                      // (possibly empty: keeps the group count uniform)
    };

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) issue(m0 + s, s);
    int buf = 0;
    for (int m = m0; m < m1; ++m) {
        issue(m + NST - 1, buf == 0 ? NST - 1 : buf - 1);
        cp_async_wait<NST - 1>();
        __syncthreads();
        const uint32_t st = sbase + (uint32_t)(buf * stage_f4) * 16u;
        float4* om = out4 + (size_t)m * mesh_out + (size_t)(r0 + lr0) * CQ + q;
        // all the staged-row reads of this thread first (independent LDS.128, no branches: padding entries
        // read the thread's piece of staged row 0 and are discarded by the selects below), then the sums in
        // storage order
        float4 v[RPT][WD];
#pragma unroll
        for (int k = 0; k < RPT; ++k)
#pragma unroll
            for (int j = 0; j < WD; ++j) v[k][j] = lds128(st + (uint32_t)(loc[k][j] < 0 ? q * 16 : loc[k][j]));
        float4 acc[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < WD; ++j) {
                const bool on = loc[k][j] >= 0;
                const float w = wv[k][j];
                const float tx = __fadd_rn(acc[k].x, __fmul_rn(v[k][j].x, w));
                const float ty = __fadd_rn(acc[k].y, __fmul_rn(v[k][j].y, w));
                const float tz = __fadd_rn(acc[k].z, __fmul_rn(v[k][j].z, w));
                const float tw = __fadd_rn(acc[k].w, __fmul_rn(v[k][j].w, w));
                acc[k].x = on ? tx : acc[k].x; acc[k].y = on ? ty : acc[k].y;
                acc[k].z = on ? tz : acc[k].z; acc[k].w = on ? tw : acc[k].w;
            }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k)
            if (r0 + lr0 + k * RSTEP < Vout) om[(size_t)k * RSTEP * CQ] = acc[k];
        __syncthreads();                                   // buffer `buf` is refilled by the next iteration's issue
        buf = buf + 1 == NST ? 0 : buf + 1;
    }
}

// ---- CSR row-sum, one output row per warp -------------------------------------------------------
// csr_rowsum_batch_kernel gives a warp 32/CQ consecutive output rows; their entry counts differ a lot for
// the transposed up-sampling matrices (1 ... 96 fine rows per coarse vertex, mean 12), so the warp runs
// max(count) iterations with most lanes idle.  Here the CQ lanes of a row piece and 32/CQ meshes share ONE
// output row (uniform trip count, entry loads broadcast), and every thread carries kPoolMeshes further
// meshes in registers.  Same per-element order of operations as the kernels above (bit-identical results).
template <int CQ>
__global__ void __launch_bounds__(256)
csr_rowsum_warp_kernel(const float* __restrict__ dy, const int* __restrict__ ptr,
                       const int* __restrict__ src, const float* __restrict__ val,
                       const float* __restrict__ gate, float* __restrict__ dx,
                       long long total_warps, int B, int Vsrc, int Vdst) {
    constexpr int ML = 32 / CQ;                            // meshes across the lanes of a warp
    constexpr int MW = ML * kPoolMeshes;                   // meshes per warp
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= total_warps) return;
    const int lane = threadIdx.x & 31, q = lane % CQ, ml = lane / CQ;
    const int k = (int)(w % Vdst);
    const int b0 = (int)(w / Vdst) * MW + ml;              // this thread's meshes: b0 + i*ML
    const int e0 = __ldg(ptr + k), e1 = __ldg(ptr + k + 1);
    const float4* dy4 = reinterpret_cast<const float4*>(dy);
    float4 acc[kPoolMeshes];
#pragma unroll
    for (int i = 0; i < kPoolMeshes; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int e = e0; e < e1; ++e) {
        const float wv = val ? __ldg(val + e) : 1.f;
        const int sr = __ldg(src + e);
        float4 v[kPoolMeshes];
#pragma unroll
        for (int i = 0; i < kPoolMeshes; ++i)
            v[i] = b0 + i * ML < B ? __ldg(dy4 + ((size_t)(b0 + i * ML) * Vsrc + sr) * CQ + q)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kPoolMeshes; ++i) {
            acc[i].x = __fadd_rn(acc[i].x, __fmul_rn(v[i].x, wv));
            acc[i].y = __fadd_rn(acc[i].y, __fmul_rn(v[i].y, wv));
            acc[i].z = __fadd_rn(acc[i].z, __fmul_rn(v[i].z, wv));
            acc[i].w = __fadd_rn(acc[i].w, __fmul_rn(v[i].w, wv));
        }
    }
#pragma unroll
    for (int i = 0; i < kPoolMeshes; ++i) {
        const int b = b0 + i * ML;
        if (b >= B) continue;
        const size_t off = ((size_t)b * Vdst + k) * CQ + q;
        float4 a = acc[i];
        if (gate) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gate) + off);
            a.x *= elu_grad_from_out(g.x); a.y *= elu_grad_from_out(g.y);
            a.z *= elu_grad_from_out(g.z); a.w *= elu_grad_from_out(g.w);
        }
        reinterpret_cast<float4*>(dx)[off] = a;
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
