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

// dW[n, s*32 + c] = sum_p part[p][(s*NO + n)*32 + c];  db[n] = sum_p part[p][J*32 + n]   (partials in CTA order)
__global__ void narrow_out_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dW,
                                         float* __restrict__ db, int S, int NO) {
    const int J = S * NO, total = J * 32 + 32;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    float t = 0.f;
    for (int p = 0; p < nparts; ++p) t += part[(size_t)p * total + idx];
    if (idx < J * 32) {
        const int j = idx >> 5, c = idx & 31;
        if (dW) dW[(size_t)(j % NO) * S * 32 + (j / NO) * 32 + c] = t;
    } else if (db && idx - J * 32 < NO) {
        db[idx - J * 32] = t;
    }
}

}  // namespace sdvae
