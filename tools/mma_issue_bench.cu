// Micro-benchmark: how fast can tcgen05.mma (kind::tf32, A from TMEM, B from shared memory) be issued / executed?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I craniofacialsd-vae_b200/csrc -o tools/mma_issue_bench tools/mma_issue_bench.cu
// One CTA per SM (persistent style, 148 CTAs or 1), one or two issuing threads (different warps).
#include <cstdio>
#include <cstdlib>
#include "spiral_conv_umma.cuh"
namespace sdvae { char g_last_error[512] = ""; }
using namespace sdvae::umma;

struct Res { long long clk; };

// mode: 0 = all N=64 ; 1 = alternate N=64 / N=32 (as the conv kernels) ; 2 = all N=32 ; 3 = all N=128 ; 4 = all N=16; 5 = N=96
__global__ void __launch_bounds__(128, 1) bench(int mode, int iters, int nthreads, int commit_every, int nacc, Res* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    if ((warp == 1 || (warp == 2 && nthreads == 2))) {
        if (elect_one()) {
            const int me = warp - 1;
            const uint64_t desc0 = smem_desc_sw128(smem_u32(smem));
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
            const uint32_t d_tmem = tb + (uint32_t)(me * 192);
            const uint32_t a_tmem = tb + 384u + (uint32_t)(me * 64);
            const long long t0 = clock64();
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo0 + 2u * k);
                    // accumulator rotation: MMA number j = 2k (+1) goes to accumulator j % nacc (32..64 columns apart)
                    const uint32_t d0 = d_tmem + (uint32_t)(((2 * k) % nacc) * 64), d1 = d_tmem + (uint32_t)(((2 * k + 1) % nacc) * 64);
                    const int n0 = mode == 0 ? 64 : mode == 1 ? 64 : mode == 2 ? 32 : mode == 3 ? 128 : mode == 4 ? 16 : 96;
                    const int n1 = mode == 1 ? 32 : n0;
                    umma_tf32_ts(d0, a_tmem + k * 8, bd, idesc_tf32(128, n0), 1u);
                    umma_tf32_ts(d1, a_tmem + 32 + k * 8, bd, idesc_tf32(128, n1), 1u);
                }
                if (commit_every > 0 && (it + 1) % commit_every == 0) { umma_commit(bar + me); mbar_wait(bar + me, ph); ph ^= 1; }
            }
            umma_commit(bar + me);
            mbar_wait(bar + me, ph);
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[me].clk = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 148;
    Res* d; cudaMalloc(&d, 2 * sizeof(Res));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    const char* names[] = {"N=64,N=64", "N=64,N=32", "N=32,N=32", "N=128,N=128", "N=16,N=16", "N=96,N=96"};
    const int tensor_clk[] = {64, 48, 32, 128, 16, 96};       // floor per pair: 128*N/256 each
    for (int nthreads = 1; nthreads <= 2; ++nthreads)
        for (int nacc : {1, 2, 3})
            for (int ce : {0, 3})
                for (int mode : {1, 0, 2}) {
                    cudaMemset(d, 0, 2 * sizeof(Res));
                    bench<<<grid, 128, 48 * 1024>>>(mode, iters, nthreads, ce, nacc, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                    Res h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                    printf("threads %d accumulators %d commit+wait every %d chunks  %-12s : %.1f clk per 32-wide chunk (4 pairs; tensor floor %d)%s\n", nthreads, nacc, ce, names[mode],
                           (double)h[0].clk / iters, 4 * tensor_clk[mode], nthreads == 2 ? "  [each of 2 threads]" : "");
                }
    return 0;
}
