// Micro-benchmark: throughput of tcgen05.st shapes (TMEM stores) with 1..16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I craniofacialsd-vae_b200/csrc -o tools/sttm_bench tools/sttm_bench.cu
#include <cstdio>
#include <cstdlib>
#include "spiral_conv_tile.cuh"
namespace sdvae { char g_last_error[512] = ""; }
using namespace sdvae::umma;
using namespace sdvae::tile;

__device__ __forceinline__ void st_16x256b_x4(uint32_t taddr, const float (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}
__device__ __forceinline__ void st_16x128b_x8(uint32_t taddr, const float (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}
__device__ __forceinline__ void st_32x32b_x16(uint32_t taddr, const float (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}

// mode 0: 32x32b.x32 (32 lanes x 32 cols = 4 KB) ; 1: 16x256b.x8 (16 lanes x 64 cols = 4 KB) ; 2: 16x256b.x4 (2 KB)
// 3: 16x128b.x8 (16 lanes x 32 cols = 2 KB) ; 4: 32x32b.x16 (2 KB)
__global__ void __launch_bounds__(576, 1) bench(int mode, int iters, int nwarps, int wait_every, int mma_warps, long long* out) {
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(1024) uint8_t bsm[16384];
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384 / 4; i += 576) reinterpret_cast<float*>(bsm)[i] = 0.001f * (i & 255);
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (float)(tid + j);
    const uint32_t t_a = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 3) * 64);
    __syncthreads();
    const long long t0 = clock64();
    if (warp >= 16) {
        // MMA-issuing warps (16, 17): D at columns 256.. / 352.., A at columns 448..: run until the store warps' work is roughly done
        if (warp - 16 < mma_warps && elect_one()) {
            const int me = warp - 16;
            const uint64_t desc0 = smem_desc_sw128(smem_u32(bsm));
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
            const uint32_t d_tmem = tb + 256u + (uint32_t)(me * 96), a_tmem = tb + 448u + (uint32_t)(me * 32);
            const long long tm0 = clock64();
            const int n_mma = iters;         // chunks of 4 MMAs
            for (int it = 0; it < n_mma; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo0 + 2u * k);
                    umma_tf32_ts(d_tmem, a_tmem + k * 8, bd, me == 0 ? idesc_tf32(128, 64) : idesc_tf32(128, 32), 1u);
                }
            }
            umma_commit(bar + me);
            mbar_wait(bar + me, 0);
            if (blockIdx.x == 0) out[1 + me] = clock64() - tm0;
        }
        __syncwarp();
    } else if (warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
            if (mode == 0) tmem_st32(t_a, v);
            else if (mode == 1) tmem_st_16x256b_x8(t_a + ((uint32_t)((it & 1) * 16) << 16), v);
            else if (mode == 2) st_16x256b_x4(t_a + ((uint32_t)((it & 1) * 16) << 16), v);
            else if (mode == 3) st_16x128b_x8(t_a + ((uint32_t)((it & 1) * 16) << 16), v);
            else st_32x32b_x16(t_a, v);
            if (wait_every > 0 && (it + 1) % wait_every == 0) tmem_st_wait();
        }
        tmem_st_wait();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 32);
    const int iters = 2000;
    const char* names[] = {"32x32b.x32 (4 KB)", "16x256b.x8 (4 KB)", "16x256b.x4 (2 KB)", "16x128b.x8 (2 KB)", "32x32b.x16 (2 KB)"};
    const int kb[] = {4, 4, 2, 2, 2};
    for (int mw : {0, 1, 2})
        for (int nw : {0, 4, 8, 16})
            for (int we : {0, 2})
                for (int mode : {1, 0}) {
                    if (nw == 0 && (we || mode == 0 || mw == 0)) continue;
                    cudaMemset(d, 0, 32);
                    bench<<<148, 576>>>(mode, iters, nw, we, mw, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
                    printf("mma warps %d | store warps %2d wait::st every %d  %-18s: ", mw, nw, we, names[mode]);
                    if (mw) printf("MMA: %.1f clk per 4-MMA chunk (hi warp)%s | ", (double)h[1] / iters, mw == 2 ? "" : "");
                    if (mw == 2) printf("lo warp %.1f | ", (double)h[2] / iters);
                    printf("whole kernel %.0f clk", (double)h[0]);
                    if (nw) printf(" = %.1f clk per store per warp if stores bound it, %.0f B/clk/SM", (double)h[0] / iters, kb[mode] * 1024.0 * iters * nw / (double)h[0]);
                    printf("\n");
                }
    return 0;
}
