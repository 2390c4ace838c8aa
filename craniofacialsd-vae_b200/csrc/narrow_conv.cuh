// Backward of the narrow-OUTPUT SpiralConv layer (32 -> 3, the last layer of the decoder, model.py:135-136,
// 172; autograd of model.py:34,40) in ONE pass, on the fp32 FMA units.
//
// With J = S*NO (= 27) and the per-vertex vector
//     G[u, s*NO + n] = sum_{v in cell(u, s)} dy[v, n]          (cell(u, s) = rows v with idx[v, s] == u)
// the three gradients of the layer are
//     dx[u, c]          = gate(u, c) * sum_j G[u, j] * W[j % NO, (j / NO)*32 + c]
//     dW[n, s*32 + c]   = sum_{b, u} G[u, s*NO + n] * x[b, u, c]
//     db[n]             = sum_{b, u} G[u, n]                  (every row v lies in exactly one cell of slot 0)
// The slot-packed path (slot_pack.cuh) materialised G as a [V, 32] tensor and ran two dense tcgen05
// contractions over it (pack 0.47 ms + dx 0.22 ms + dW 0.45 ms at 256 meshes: G written once and read twice,
// x read twice).  Here G never leaves the SM: the mesh's dy (204 KB for the craniofacial template) is resident
// in shared memory, lane j of a warp sums its cell (u, j/NO) for channel j%NO, the 27 values go through a
// warp-private shared-memory row to every lane, and lane c (= input channel) does the 27 + 27 FMAs of dx[u, c]
// and dW[., c] against weights / accumulators it keeps in registers.  HBM traffic: dy + x in, dx out.
// Deterministic: fixed summation orders, per-CTA partials reduced in CTA order by narrow_out_reduce_kernel.
#pragma once
#include "common.cuh"
#include "pool_misc.cuh"

namespace sdvae {

constexpr int kNarrowThreads = 512;
constexpr int kNarrowWarps = kNarrowThreads / 32;
constexpr int kNarrowRows = 4;                       // vertices in flight per warp

template <int S, int NO>
struct NarrowCfg {
    static constexpr int J = S * NO;
    static constexpr int PART = J * 32 + 32;         // floats per CTA partial: dW image [J][32] + column sums of G
    // shared memory (floats): [max(4 + R*NO, kNarrowWarps * PART)] dy of the mesh / final reduction, then the
    // warp-private G rows
    static size_t main_floats(int R) {
        const size_t a = 8 + (size_t)R * NO, b = (size_t)kNarrowWarps * PART;     // 4 phase + mesh + zero pad
        return ((a > b ? a : b) + 3) & ~(size_t)3;
    }
    static size_t smem_bytes(int R) { return (main_floats(R) + (size_t)kNarrowWarps * kNarrowRows * 32) * 4; }
};

template <int S, int NO>
__global__ void __launch_bounds__(kNarrowThreads, 1)
narrow_out_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                      const int* __restrict__ cell_ptr, const int* __restrict__ cell_src,
                      const int* __restrict__ cell_pack, const float* __restrict__ W,
                      float* __restrict__ dx, float* __restrict__ part,
                      int B, int parts, int R, int Vin, int gated, int main_floats) {
    using Cfg = NarrowCfg<S, NO>;
    constexpr int J = Cfg::J, JQ = (J + 3) / 4;
    extern __shared__ float nb_smem[];
    float* xs_raw = nb_smem;                                         // [4 + R*NO + 4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* gsw = nb_smem + main_floats + warp * kNarrowRows * 32;    // this warp's G rows
    const bool live = lane < J;
    const int s = live ? lane / NO : 0, c = live ? lane - s * NO : 0;    // gather role: lane j = (slot, channel of dy)
    float w[J], accw[J];                                                 // contraction role: lane = channel of x / dx
#pragma unroll
    for (int j = 0; j < J; ++j) {
        w[j] = __ldg(W + (size_t)(j % NO) * S * 32 + (j / NO) * 32 + lane);
        accw[j] = 0.f;
    }
    float gsum = 0.f;
    const int n = R * NO;
    const int rows_per_part = (Vin + parts - 1) / parts;
    for (int item = blockIdx.x; item < B * parts; item += gridDim.x) {
        const int b = item / parts, pi = item - b * parts;
        const float* src = dy + (size_t)b * n;
        // 3-channel rows: a mesh starts at any 4-byte phase; keep the source's 16-byte phase in shared memory
        // (as slot_pack_smem_kernel does) so that all but <= 6 elements move as 16-byte cp.async
        const int phase = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
        float* xs = xs_raw + phase;
        const int head = (4 - phase) & 3;
        const int n4 = (n - head) >> 2, tail0 = head + 4 * n4;
        __syncthreads();                                             // previous item's reads are done
        for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async16(xs + head + 4 * i, src + head + 4 * i);
        if ((int)threadIdx.x < head) cp_async4(xs + threadIdx.x, src + threadIdx.x);
        if (tail0 + (int)threadIdx.x < n) cp_async4(xs + tail0 + threadIdx.x, src + tail0 + threadIdx.x);
        cp_async_commit();
        if ((int)threadIdx.x < NO) xs[n + threadIdx.x] = 0.f;         // zero pad behind the mesh: the "no row" offset
        cp_async_wait<0>();
        __syncthreads();
        const int r_begin = pi * rows_per_part;
        const int r_end = min(Vin, r_begin + rows_per_part);
        const float* xb = x + (size_t)b * Vin * 32 + lane;
        float* dxb = dx ? dx + (size_t)b * Vin * 32 + lane : nullptr;
        // Gather: lane j sums its cell (u, j/NO).  The cell's first four source rows come packed in ONE 8-byte
        // word (cell_pack) as 16-bit ELEMENT OFFSETS row*NO into the mesh's dy; an absent row is the offset R*NO of
        // a zero pad behind the mesh (adding +0 needs no compare / select), and 0xFFFF in the fourth field marks a
        // longer cell whose rows 4.. are read from the CSR (rare: 42 of 153 351 cells at level 0 of the
        // craniofacial template).  No dependent global load in the common case, ~4 instructions per entry.
        const int step = kNarrowWarps * kNarrowRows;
        const uint2* cpk = reinterpret_cast<const uint2*>(cell_pack);
        auto load_pack = [&](int r0, uint2 (&pk)[kNarrowRows]) {
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) pk[i] = __ldg(cpk + (size_t)min(r0 + i, Vin - 1) * S + s);
        };
        // prefetch distance TWO iterations: the compiler sinks a load to just before the back-edge of the iteration
        // that issues it, so words fetched for the next iteration arrived a few instructions ahead of their use
        // (15 % of all stall samples); fetched for the one after, they have a whole iteration to land
        uint2 nxt[kNarrowRows], nxt2[kNarrowRows];
        const int r_first = r_begin + warp * kNarrowRows;
        load_pack(r_first, nxt);
        load_pack(r_first + step, nxt2);
        const float* xsc = xs + c;
        for (int r0 = r_first; r0 < r_end; r0 += step) {
            uint2 pk[kNarrowRows];
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) { pk[i] = nxt[i]; nxt[i] = nxt2[i]; }
            load_pack(r0 + 2 * step, nxt2);
            float xv[kNarrowRows];
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) xv[i] = __ldg(xb + (size_t)min(r0 + i, Vin - 1) * 32);
            float acc[kNarrowRows];
            bool any_long = false;
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) {
                const uint32_t o3 = pk[i].y >> 16;
                const bool lng = o3 == 0xffffu;
                any_long |= lng;
                const float v0 = xsc[pk[i].x & 0xffffu], v1 = xsc[pk[i].x >> 16];
                const float v2 = xsc[pk[i].y & 0xffffu], v3 = xsc[lng ? (uint32_t)n : o3];
                acc[i] = __fadd_rn(__fadd_rn(__fadd_rn(v0, v1), v2), v3);      // stored (ascending-row) order
            }
            if (__any_sync(0xffffffffu, any_long)) {
#pragma unroll
                for (int i = 0; i < kNarrowRows; ++i) {
                    if ((pk[i].y >> 16) == 0xffffu) {                // long cell: rows 4.. from the CSR
                        const int u = min(r0 + i, Vin - 1);
                        const int e1 = __ldg(cell_ptr + (size_t)u * S + s + 1);
                        for (int e = __ldg(cell_ptr + (size_t)u * S + s) + 3; e < e1; ++e)
                            acc[i] = __fadd_rn(acc[i], xsc[__ldg(cell_src + e) * NO]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) {
                // dead lanes and rows past the range contribute zeros (their dx store is predicated off below)
                acc[i] = (live && r0 + i < r_end) ? acc[i] : 0.f;
                gsum += acc[i];
                gsw[i * 32 + lane] = acc[i];
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) {
                float g[4 * JQ];
#pragma unroll
                for (int k = 0; k < JQ; ++k) {
                    const float4 t = reinterpret_cast<const float4*>(gsw + i * 32)[k];
                    g[4 * k] = t.x; g[4 * k + 1] = t.y; g[4 * k + 2] = t.z; g[4 * k + 3] = t.w;
                }
                float d0 = 0.f, d1 = 0.f, d2 = 0.f;                  // three interleaved chains, fixed order
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (j % 3 == 0) d0 = fmaf(g[j], w[j], d0);
                    else if (j % 3 == 1) d1 = fmaf(g[j], w[j], d1);
                    else d2 = fmaf(g[j], w[j], d2);
                }
                float d = (d0 + d1) + d2;
                if (gated) d *= elu_grad_from_out(xv[i]);
                if (dxb && r0 + i < r_end) dxb[(size_t)(r0 + i) * 32] = d;
#pragma unroll
                for (int j = 0; j < J; ++j) accw[j] = fmaf(g[j], xv[i], accw[j]);     // g = 0 for rows past the range
            }
            __syncwarp();                                            // the G rows are rewritten by the next iteration
        }
    }
    // per-CTA partial: warps summed in warp order
    __syncthreads();
    float* red = nb_smem;                                            // [kNarrowWarps][J*32] | [kNarrowWarps][32]
#pragma unroll
    for (int j = 0; j < J; ++j) red[(warp * J + j) * 32 + lane] = accw[j];
    red[kNarrowWarps * J * 32 + warp * 32 + lane] = gsum;
    __syncthreads();
    float* po = part + (size_t)blockIdx.x * Cfg::PART;
    for (int idx = threadIdx.x; idx < Cfg::PART; idx += blockDim.x) {
        float t = 0.f;
        if (idx < J * 32) {
            for (int wq = 0; wq < kNarrowWarps; ++wq) t += red[wq * J * 32 + idx];
        } else {
            for (int wq = 0; wq < kNarrowWarps; ++wq) t += red[kNarrowWarps * J * 32 + wq * 32 + (idx - J * 32)];
        }
        po[idx] = t;
    }
}

// dW[n, s*32 + c] = sum_p part[p][(s*NO + n)*32 + c];  db[n] = sum_p part[p][J*32 + n].  One warp per output element:
// lane l adds the partials p = l, l+32, ... in order, a fixed shuffle tree adds the lanes (deterministic; the
// one-thread-per-element version walked <= 148 dependent loads and took 30 us).
__device__ __forceinline__ float narrow_partial_sum(const float* __restrict__ part, int nparts, int total, int idx) {
    const int lane = threadIdx.x & 31;
    float t = 0.f;
    for (int p = lane; p < nparts; p += 32) t += part[(size_t)p * total + idx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

__global__ void narrow_out_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dW,
                                         float* __restrict__ db, int S, int NO) {
    const int J = S * NO, total = J * 32 + 32;
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (idx >= total) return;
    const float t = narrow_partial_sum(part, nparts, total, idx);
    if ((threadIdx.x & 31) != 0) return;
    if (idx < J * 32) {
        const int j = idx >> 5, c = idx & 31;
        if (dW) dW[(size_t)(j % NO) * S * 32 + (j / NO) * 32 + c] = t;
    } else if (db && idx - J * 32 < NO) {
        db[idx - J * 32] = t;
    }
}


// ---- narrow-output layer forward (32 -> 3) ------------------------------------------------------------
//   y[b, v, n] = bias[n] + sum_{s, c} W[n, s*32 + c] * x[b, idx[v, s], c]                      (model.py:27-41)
// On the tensor-core kernel this layer cost as much as a 32 -> 32 layer (it is bound by the gather of 9 x 128
// bytes per vertex through L2, for 3 outputs).  Here the DISTINCT source rows of a tile of kNarrowTile output
// rows are staged once per mesh in shared memory (the plan and the cp.async ring of pool_ell_fwd_staged_kernel:
// 3.5 rows per output row instead of 9), the 8 lanes of an output row read whole staged rows (conflict-free
// LDS.128, 4 channels each) and do the 9 x 4 x 3 FMAs against weights held in registers; a 3-step shuffle tree
// adds the 8 partial sums.  Deterministic (fixed orders).
constexpr int kNarrowFwdThreads = 256;
constexpr int kNarrowTile = 64;           // output rows per tile (= tables.NARROW_FWD_TILE)
constexpr int kNarrowMinStages = 3, kNarrowMaxStages = 5;   // cp.async ring depth: as deep as shared memory allows
constexpr int kNarrowMaxIssue = 12;       // cp.async per thread and mesh: ucap * 8 <= 12 * 256

template <int S, int NO, int NST>
__global__ void __launch_bounds__(kNarrowFwdThreads, 1)
narrow_out_fwd_kernel(const float* __restrict__ x, const int* __restrict__ tile_ptr,
                      const int* __restrict__ stage_src, const int* __restrict__ loc_tab,
                      const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ out,
                      int B, int Vin, int Vout, int L, int ucap, int MG) {
    constexpr int CQ = 8;                                        // 16-byte pieces of a 32-channel row
    constexpr int PASSES = kNarrowTile * CQ / kNarrowFwdThreads; // output rows per thread and mesh
    constexpr int RSTEP = kNarrowFwdThreads / CQ;
    extern __shared__ float4 nf_stage[];                         // [NST][ucap * CQ]
    const int tile = (int)blockIdx.x % L, grp = (int)blockIdx.x / L;
    const int m0 = grp * MG, m1 = min(B, m0 + MG);
    if (m0 >= m1) return;
    const int u0 = __ldg(tile_ptr + tile), U = __ldg(tile_ptr + tile + 1) - u0;
    const int r0 = tile * kNarrowTile;
    const int tid = threadIdx.x, q = tid % CQ, lr0 = tid / CQ;
    const size_t mesh_in = (size_t)Vin * CQ;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int stage_f4 = ucap * CQ;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(nf_stage);

    int src_off[kNarrowMaxIssue];
#pragma unroll
    for (int i = 0; i < kNarrowMaxIssue; ++i) {
        const int p = tid + i * kNarrowFwdThreads;
        src_off[i] = p < U * CQ ? __ldg(stage_src + u0 + p / CQ) * CQ + q : -1;
    }
    uint32_t loc[PASSES][S];                                     // byte offset of (staged row, piece q) in a stage
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
        const int r = min(r0 + lr0 + k * RSTEP, Vout - 1);       // rows past the end shadow the last row (not stored)
#pragma unroll
        for (int s = 0; s < S; ++s) loc[k][s] = (uint32_t)(__ldg(loc_tab + (size_t)r * S + s) * CQ + q) * 16u;
    }
    float w[S][4][NO];                                           // W[n, s*32 + 4q + i]
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int n = 0; n < NO; ++n) w[s][i][n] = __ldg(W + (size_t)n * S * 32 + s * 32 + 4 * q + i);
    const float bq = (bias != nullptr && q < NO) ? __ldg(bias + q) : 0.f;

    auto issue = [&](int m, int buf) {
        if (m < m1) {
            const float4* xm = x4 + (size_t)m * mesh_in;
            float4* dst = nf_stage + (size_t)buf * stage_f4 + tid;
#pragma unroll
            for (int i = 0; i < kNarrowMaxIssue; ++i)
                if (src_off[i] >= 0) cp_async16(dst + i * kNarrowFwdThreads, xm + src_off[i]);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int st = 0; st < NST - 1; ++st) issue(m0 + st, st);
    int buf = 0;
    for (int m = m0; m < m1; ++m) {
        issue(m + NST - 1, buf == 0 ? NST - 1 : buf - 1);
        cp_async_wait<NST - 1>();
        __syncthreads();
        const uint32_t st = sbase + (uint32_t)(buf * stage_f4) * 16u;
        float acc[PASSES][NO];
#pragma unroll
        for (int k = 0; k < PASSES; ++k)
#pragma unroll
            for (int n = 0; n < NO; ++n) acc[k][n] = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int k = 0; k < PASSES; ++k) {
                const float4 v = lds128(st + loc[k][s]);
#pragma unroll
                for (int n = 0; n < NO; ++n) {
                    acc[k][n] = fmaf(v.x, w[s][0][n], acc[k][n]);
                    acc[k][n] = fmaf(v.y, w[s][1][n], acc[k][n]);
                    acc[k][n] = fmaf(v.z, w[s][2][n], acc[k][n]);
                    acc[k][n] = fmaf(v.w, w[s][3][n], acc[k][n]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PASSES; ++k) {
#pragma unroll
            for (int n = 0; n < NO; ++n) {
                float a = acc[k][n];
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                a += __shfl_xor_sync(0xffffffffu, a, 4);
                acc[k][n] = a;
            }
            const int r = r0 + lr0 + k * RSTEP;
            if (q < NO && r < Vout) {
                float o = acc[k][0];
#pragma unroll
                for (int n = 1; n < NO; ++n) o = q == n ? acc[k][n] : o;
                out[((size_t)m * Vout + r) * NO + q] = o + bq;
            }
        }
        __syncthreads();
        buf = buf + 1 == NST ? 0 : buf + 1;
    }
}


// ---- narrow-INPUT layer (3 -> 32, the first encoder block, model.py:104-110) ----------------------------
//   forward      y[b, r, o] = act( bias[o] + sum_j P[b, r, j] * W[o, j] ),  P[b, r, s*CI + c] = x[b, idx[r, s], c]
//   weight grad  dW[o, j] = sum_{b, r} dpre[b, r, o] * P[b, r, j],   db[o] = sum_{b, r} dpre[b, r, o]
// with the mesh's whole input ([Vin, 3]: 204 KB) resident in shared memory, as in narrow_out_bwd_kernel: lane j
// gathers its value of P (one coalesced index load, one LDS), the J = 27 values reach every lane through a
// warp-private shared-memory row, and lane o (= output channel) does the 27 FMAs against the weights (forward)
// or accumulators (weight gradient) it keeps in registers.  P is never written to memory (the slot-packed path
// wrote it in the forward pass and read it twice).  idx may be a row-restricted table (R kept rows).
template <int S, int CI>
struct NarrowInCfg {
    static constexpr int J = S * CI;
    static constexpr int PART = J * 32 + 32;
    static size_t main_floats(int Vin) {
        const size_t a = 8 + (size_t)Vin * CI, b = (size_t)kNarrowWarps * PART;
        return ((a > b ? a : b) + 3) & ~(size_t)3;
    }
    static size_t smem_bytes(int Vin) { return (main_floats(Vin) + (size_t)kNarrowWarps * kNarrowRows * 32) * 4; }
};

// stage mesh `src` ([n] floats at any 4-byte phase) behind xs_raw, keeping the source's 16-byte phase; returns xs
__device__ __forceinline__ float* narrow_stage_mesh(float* xs_raw, const float* src, int n) {
    const int phase = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
    float* xs = xs_raw + phase;
    const int head = (4 - phase) & 3;
    const int n4 = (n - head) >> 2, tail0 = head + 4 * n4;
    __syncthreads();                                                 // previous item's reads are done
    for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async16(xs + head + 4 * i, src + head + 4 * i);
    if ((int)threadIdx.x < head) cp_async4(xs + threadIdx.x, src + threadIdx.x);
    if (tail0 + (int)threadIdx.x < n) cp_async4(xs + tail0 + threadIdx.x, src + tail0 + threadIdx.x);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    return xs;
}

// MODE 0: forward (out = y, aux = bias, act);  MODE 1: weight gradient (aux = dpre, out = per-CTA partials)
template <int S, int CI, int MODE>
__global__ void __launch_bounds__(kNarrowThreads, 1)
narrow_in_kernel(const float* __restrict__ x, const int* __restrict__ idx, const float* __restrict__ W,
                 const float* __restrict__ aux, float* __restrict__ out, int B, int parts, int R, int Vin,
                 int act, int main_floats) {
    using Cfg = NarrowInCfg<S, CI>;
    constexpr int J = Cfg::J, JQ = (J + 3) / 4;
    extern __shared__ float nb_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* gsw = nb_smem + main_floats + warp * kNarrowRows * 32;
    const bool live = lane < J;
    const int s = live ? lane / CI : 0, c = live ? lane - s * CI : 0;
    float w[J];                                                      // MODE 0: W[lane, j];  MODE 1: dW accumulators
#pragma unroll
    for (int j = 0; j < J; ++j) w[j] = MODE == 0 ? __ldg(W + (size_t)lane * J + j) : 0.f;
    float bsum = MODE == 0 ? (aux ? __ldg(aux + lane) : 0.f) : 0.f;  // MODE 0: bias[lane];  MODE 1: db accumulator
    const int n = Vin * CI;
    const int rows_per_part = (R + parts - 1) / parts;
    const int step = kNarrowWarps * kNarrowRows;
    for (int item = blockIdx.x; item < B * parts; item += gridDim.x) {
        const int b = item / parts, pi = item - b * parts;
        const float* xs = narrow_stage_mesh(nb_smem, x + (size_t)b * n, n) + c;
        const int r_begin = pi * rows_per_part;
        const int r_end = min(R, r_begin + rows_per_part);
        const float* ab = MODE == 1 ? aux + (size_t)b * R * 32 + lane : nullptr;
        float* ob = MODE == 0 ? out + (size_t)b * R * 32 + lane : nullptr;
        auto load_idx = [&](int r0, int (&v)[kNarrowRows]) {
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) v[i] = __ldg(idx + (size_t)min(r0 + i, R - 1) * S + s);
        };
        int nxt[kNarrowRows], nxt2[kNarrowRows];                     // two iterations ahead (see narrow_out_bwd_kernel)
        const int r_first = r_begin + warp * kNarrowRows;
        load_idx(r_first, nxt);
        load_idx(r_first + step, nxt2);
        for (int r0 = r_first; r0 < r_end; r0 += step) {
            int v[kNarrowRows];
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) { v[i] = nxt[i]; nxt[i] = nxt2[i]; }
            load_idx(r0 + 2 * step, nxt2);
            float dv[kNarrowRows];
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < kNarrowRows; ++i)
                    dv[i] = r0 + i < r_end ? __ldg(ab + (size_t)min(r0 + i, R - 1) * 32) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) gsw[i * 32 + lane] = live ? xs[v[i] * CI] : 0.f;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < kNarrowRows; ++i) {
                float g[4 * JQ];
#pragma unroll
                for (int k = 0; k < JQ; ++k) {
                    const float4 t = reinterpret_cast<const float4*>(gsw + i * 32)[k];
                    g[4 * k] = t.x; g[4 * k + 1] = t.y; g[4 * k + 2] = t.z; g[4 * k + 3] = t.w;
                }
                if (MODE == 0) {
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f;              // three interleaved chains, fixed order
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        if (j % 3 == 0) d0 = fmaf(g[j], w[j], d0);
                        else if (j % 3 == 1) d1 = fmaf(g[j], w[j], d1);
                        else d2 = fmaf(g[j], w[j], d2);
                    }
                    float d = ((d0 + d1) + d2) + bsum;
                    if (act == 1) d = elu_fast(d);                    // SDVAE_ACT_ELU
                    if (r0 + i < r_end) ob[(size_t)(r0 + i) * 32] = d;
                } else {
                    bsum += dv[i];                                   // dv = 0 for rows past the range
#pragma unroll
                    for (int j = 0; j < J; ++j) w[j] = fmaf(g[j], dv[i], w[j]);
                }
            }
            __syncwarp();
        }
    }
    if (MODE == 1) {
        __syncthreads();
        float* red = nb_smem;                                        // [kNarrowWarps][J*32] | [kNarrowWarps][32]
#pragma unroll
        for (int j = 0; j < J; ++j) red[(warp * J + j) * 32 + lane] = w[j];
        red[kNarrowWarps * J * 32 + warp * 32 + lane] = bsum;
        __syncthreads();
        float* po = out + (size_t)blockIdx.x * Cfg::PART;
        for (int i = threadIdx.x; i < Cfg::PART; i += blockDim.x) {
            float t = 0.f;
            if (i < J * 32) {
                for (int wq = 0; wq < kNarrowWarps; ++wq) t += red[wq * J * 32 + i];
            } else {
                for (int wq = 0; wq < kNarrowWarps; ++wq) t += red[kNarrowWarps * J * 32 + wq * 32 + (i - J * 32)];
            }
            po[i] = t;
        }
    }
}

// dW[o, j] = sum_p part[p][j*32 + o];  db[o] = sum_p part[p][J*32 + o]      (one warp per element, as above)
__global__ void narrow_in_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dW,
                                        float* __restrict__ db, int J) {
    const int total = J * 32 + 32;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= total) return;
    const float t = narrow_partial_sum(part, nparts, total, i);
    if ((threadIdx.x & 31) != 0) return;
    if (i < J * 32) { if (dW) dW[(size_t)(i & 31) * J + (i >> 5)] = t; }
    else if (db) db[i - J * 32] = t;
}

}  // namespace sdvae
