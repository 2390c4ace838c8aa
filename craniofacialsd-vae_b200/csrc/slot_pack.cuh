// Narrow-channel SpiralConv layers (C = 3: the first encoder block and the output layer of
// model.py:104-136) through "slot packing".
//
// For a layer whose narrow side has C channels with S*C <= 32, the gathered operand of model.py:34
// has only S*C (= 27) columns per vertex.  Materialising THAT (one 128-byte row per vertex, the same
// size as a 32-channel activation row) is cheap, and it turns every pass of the layer into a DENSE
// 32 x 32 contraction on the tcgen05 kernels with an identity tile plan (S = 1):
//
//   forward, narrow input    P[r, s*C + c] = x[idx[r, s], c]              y  = elu(P Wd^T + b)
//   weight grad, narrow input                                             dW = dy^T P           (no gather of x)
//   backward, narrow output  G[u, s*C + n] = sum_{v in cell(u,s)} dy[v,n] dx = (G Wd'^T) * elu'  (Wd'[c, j] = W[n, s*32+c])
//   weight grad, narrow output                                            dW[n, s*32+c] = (G^T x)[s*C+n, c]
//
// instead of three gather-bound kernels that each move 9 x 128 bytes per vertex for 3 useful channels.
#pragma once
#include "common.cuh"

namespace sdvae {

// out[b, r, s*C + c] = sum_{e in cell(r, s)} in[b, src[e], c];  columns S*C .. 31 are zero.
// cell_ptr == nullptr: cell(r, s) = { r*S + s } (forward table, cell_src = idx[R, S]).
// One warp per output row (128-byte coalesced store); rows are summed in storage order.  A row is a chain of
// three dependent loads (cell range -> source row -> value), so each warp keeps kSlotRows rows in flight.
constexpr int kSlotRows = 4;
__global__ void slot_pack_kernel(const float* __restrict__ in, const int* __restrict__ cell_ptr,
                                 const int* __restrict__ cell_src, float* __restrict__ out,
                                 long long rows_total, int R, int Vin, int S, int C) {
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int s = lane / C, c = lane - s * C;
    const bool live = lane < S * C;
    for (long long m0 = w0 * kSlotRows; m0 < rows_total; m0 += nw * kSlotRows) {
        int e0[kSlotRows], e1[kSlotRows];
        const float* src[kSlotRows];
#pragma unroll
        for (int i = 0; i < kSlotRows; ++i) {
            const long long m = m0 + i;
            e0[i] = e1[i] = 0;
            src[i] = in;
            if (live && m < rows_total) {
                const long long b = m / R;
                const int r = (int)(m - b * R);
                src[i] = in + (size_t)b * Vin * C + c;
                if (cell_ptr == nullptr) { e0[i] = r * S + s; e1[i] = e0[i] + 1; }
                else { e0[i] = __ldg(cell_ptr + (size_t)r * S + s); e1[i] = __ldg(cell_ptr + (size_t)r * S + s + 1); }
            }
        }
        int v0[kSlotRows];
#pragma unroll
        for (int i = 0; i < kSlotRows; ++i) v0[i] = e0[i] < e1[i] ? __ldg(cell_src + e0[i]) : -1;
        float acc[kSlotRows];
#pragma unroll
        for (int i = 0; i < kSlotRows; ++i) acc[i] = v0[i] >= 0 ? __ldg(src[i] + (size_t)v0[i] * C) : 0.f;
#pragma unroll
        for (int i = 0; i < kSlotRows; ++i)                       // cells with more than one row (inverse tables)
            for (int e = e0[i] + 1; e < e1[i]; ++e) acc[i] = __fadd_rn(acc[i], __ldg(src[i] + (size_t)__ldg(cell_src + e) * C));
#pragma unroll
        for (int i = 0; i < kSlotRows; ++i)
            if (m0 + i < rows_total) out[(size_t)(m0 + i) * 32 + lane] = acc[i];
    }
}

// The same operation with the whole narrow input of one mesh ([Vin, C], 204 KB for the craniofacial
// template) resident in shared memory: the 12-byte gathers then cost shared-memory wavefronts instead of
// one 32-byte L2 sector each (the global version moves ~5x its useful bytes and ran at 0.7 ms for 256
// meshes; this one is bound by the 128-byte row stores).  Work item = (mesh, row range); every item
// loads its mesh.  1024 threads, one warp per output row, kSlotRows rows in flight per warp.
__global__ void __launch_bounds__(1024, 1)
slot_pack_smem_kernel(const float* __restrict__ in, const int* __restrict__ cell_ptr,
                      const int* __restrict__ cell_src, float* __restrict__ out,
                      int B, int parts, int R, int Vin, int S, int C) {
    extern __shared__ float xs_raw[];                           // [4 + Vin * C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = lane < S * C;
    const int s = live ? lane / C : 0, c = live ? lane - s * C : 0;     // dead lanes shadow (slot 0, channel 0)
    const int n = Vin * C;
    const int rows_per_part = (R + parts - 1) / parts;
    for (int item = blockIdx.x; item < B * parts; item += gridDim.x) {
        const int b = item / parts, part = item - b * parts;
        const float* src = in + (size_t)b * n;
        // a mesh of 3-channel rows starts at any 4-byte phase: keep the 16-byte phase of the source in shared
        // memory (element i at xs[i], xs = xs_raw + phase) so that all but <= 3 + 3 elements move as 16-byte
        // cp.async; the rest as 4-byte cp.async.  Nothing on this path waits per element.
        const int phase = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
        float* xs = xs_raw + phase;
        const int head = (4 - phase) & 3;                        // elements before the first aligned 16 bytes
        const int n4 = (n - head) >> 2, tail0 = head + 4 * n4;
        __syncthreads();                                        // previous item's reads are done
        for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async16(xs + head + 4 * i, src + head + 4 * i);
        if (threadIdx.x < head) cp_async4(xs + threadIdx.x, src + threadIdx.x);
        if (tail0 + (int)threadIdx.x < n) cp_async4(xs + tail0 + threadIdx.x, src + tail0 + threadIdx.x);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        const int r_begin = part * rows_per_part;
        const int r_end = min(R, r_begin + rows_per_part);
        float* o = out + (size_t)b * R * 32;
        // A row is a chain cell range -> first source row -> shared-memory value.  Every load is UNCONDITIONAL
        // on a clamped index (a conditional load compiles to a branch that waits for its predicate's data, which
        // serialised the rows of a warp: 5 us per 128 rows), and the two index loads run two and one iterations
        // ahead of their use.
        const int step = 32 * kSlotRows;
        auto load_range = [&](int r0, int (&e0)[kSlotRows], int (&e1)[kSlotRows]) {
#pragma unroll
            for (int i = 0; i < kSlotRows; ++i) {
                const int r = min(r0 + i, R - 1);
                if (cell_ptr == nullptr) { e0[i] = r * S + s; e1[i] = e0[i] + 1; }
                else { e0[i] = __ldg(cell_ptr + (size_t)r * S + s); e1[i] = __ldg(cell_ptr + (size_t)r * S + s + 1); }
            }
        };
        auto load_first = [&](const int (&e0)[kSlotRows], const int (&e1)[kSlotRows], int (&v0)[kSlotRows]) {
#pragma unroll
            for (int i = 0; i < kSlotRows; ++i)                 // an empty cell reads the entry before it (unused)
                v0[i] = __ldg(cell_src + (e0[i] < e1[i] ? e0[i] : max(e0[i] - 1, 0)));
        };
        int a0[kSlotRows], a1[kSlotRows];                       // ranges of the iteration after next
        int b0[kSlotRows], b1[kSlotRows], bv[kSlotRows];        // ranges + first source row of the next iteration
        const int r_first = r_begin + warp * kSlotRows;
        load_range(r_first, b0, b1);
        load_first(b0, b1, bv);
        load_range(r_first + step, a0, a1);
        for (int r0 = r_first; r0 < r_end; r0 += step) {
            int e0[kSlotRows], e1[kSlotRows], v0[kSlotRows];
#pragma unroll
            for (int i = 0; i < kSlotRows; ++i) { e0[i] = b0[i]; e1[i] = b1[i]; v0[i] = bv[i]; b0[i] = a0[i]; b1[i] = a1[i]; }
            load_first(b0, b1, bv);
            load_range(r0 + 2 * step, a0, a1);
            float acc[kSlotRows];
#pragma unroll
            for (int i = 0; i < kSlotRows; ++i) {
                const float v = xs[v0[i] * C + c];
                acc[i] = (live && e0[i] < e1[i]) ? v : 0.f;
            }
#pragma unroll
            for (int i = 0; i < kSlotRows; ++i)                   // cells with more than one row (inverse tables)
                if (live)
                    for (int e = e0[i] + 1; e < e1[i]; ++e) acc[i] = __fadd_rn(acc[i], xs[__ldg(cell_src + e) * C + c]);
#pragma unroll
            for (int i = 0; i < kSlotRows; ++i)
                if (r0 + i < r_end) o[(size_t)(r0 + i) * 32 + lane] = acc[i];
        }
    }
}

// Dense 32 x 32 weight of a slot-packed layer (rows n, columns k), from the layer's own weight:
//   mode 0 (narrow input,  W [N, S*C])     Wd[n, k] = k < S*C ? W[n, k] : 0                      (n < N)
//   mode 1 (narrow output, W [C, S*32])    Wd[c, j] = j < S*C ? W[j % C, (j / C)*32 + c] : 0
__global__ void slot_weight_kernel(const float* __restrict__ W, float* __restrict__ Wd, int mode, int N, int S, int C) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 1024) return;
    const int row = t >> 5, col = t & 31;
    float v = 0.f;
    if (col < S * C) {
        if (mode == 0) { if (row < N) v = W[(size_t)row * S * C + col]; }
        else v = W[(size_t)(col % C) * S * 32 + (col / C) * 32 + row];
    }
    Wd[t] = v;
}

// Scatter the dense 32 x 32 weight gradient (and bias gradient) of a slot-packed layer back:
//   mode 0   dW[n, k] = dWd[n, k] (k < S*C, n < N),           db[n] = dbd[n]
//   mode 1   dW[n, s*32 + c] = dWd[s*C + n, c] (n < C),       db[n] = dbd[n]   (slot 0 is the vertex itself)
__global__ void slot_grad_kernel(const float* __restrict__ dWd, const float* __restrict__ dbd,
                                 float* __restrict__ dW, float* __restrict__ db, int mode, int N, int S, int C) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 1024) return;
    const int row = t >> 5, col = t & 31;
    if (mode == 0) {
        if (row < N && col < S * C) dW[(size_t)row * S * C + col] = dWd[t];
        if (db && col == 0 && row < N) db[row] = dbd[row];
    } else {
        if (row < S * C) dW[(size_t)(row % C) * S * 32 + (row / C) * 32 + col] = dWd[t];
        if (db && col == 0 && row < C) db[row] = dbd[row];
    }
}

}  // namespace sdvae
