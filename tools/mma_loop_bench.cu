// Micro-benchmark of the MMA-issuing thread's per-chunk loop: 4 x (N=64, N=32) kind::tf32 MMAs (A from TMEM, B from shared
// memory) + optional tcgen05.commit per `ce` chunks + optional ld.acquire poll of an (already satisfied) shared counter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I craniofacialsd-vae_b200/csrc -o tools/mma_loop_bench tools/mma_loop_bench.cu
#include <cstdio>
#include <cstdlib>
#include "spiral_conv_tile.cuh"
namespace sdvae { char g_last_error[512] = ""; }
using namespace sdvae::umma;
using namespace sdvae::tile;

__device__ __forceinline__ uint32_t ld_acquire_a(uint32_t addr) {   // (lived in spiral_conv_tile.cuh while the counter hand-off variant did)
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_a(uint32_t addr) {
    uint32_t v; asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ uint32_t ld_volatile_a(uint32_t addr) {
    uint32_t v; asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ bool try_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity));
    return ok != 0;
}
__device__ __forceinline__ bool test_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity));
    return ok != 0;
}
__device__ __forceinline__ bool try_wait_acq(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// POLL: 0 none, 1 ld.acquire, 2 ld.relaxed, 3 ld.volatile, 4 mbarrier.try_wait.relaxed (phase already complete), 5 mbarrier.test_wait.relaxed,
//       6 mbarrier.try_wait (acquire)
template <int CE, int POLL, bool FENCE>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
    __shared__ __align__(1024) uint8_t bsm[16384];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t cnt[8];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384 / 4; i += 128) reinterpret_cast<float*>(bsm)[i] = 0.001f * (i & 255);
    if (tid == 0) { for (int i = 0; i < 8; ++i) { mbar_init(bar + i, 1); cnt[i] = 0x7fffffffu; } fence_barrier_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    if (warp == 1) {
        if (elect_one()) {
            const uint64_t desc0 = smem_desc_sw128(smem_u32(bsm));
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
            const uint32_t cnt_a = smem_u32(cnt);
            int st = 0; uint32_t need = 4u;
            const long long t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
                if (POLL == 1) { while (ld_acquire_a(cnt_a + (uint32_t)st * 4u) < need) {} }
                if (POLL == 2) { while (ld_relaxed_a(cnt_a + (uint32_t)st * 4u) < need) {} }
                if (POLL == 3) { while (ld_volatile_a(cnt_a + (uint32_t)st * 4u) < need) {} }
                if (POLL == 4) { while (!try_wait_relaxed(smem_u32(bar + 6), 1u)) {} }
                if (POLL == 5) { while (!test_wait_relaxed(smem_u32(bar + 6), 1u)) {} }
                if (POLL == 6) { while (!try_wait_acq(smem_u32(bar + 6), 1u)) {} }
                if (FENCE) tc_fence_after();
                const uint32_t a_hi = tb + (uint32_t)(128 + st * 64), a_lo = a_hi + 32;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo0 + 2u * k);
                    umma_tf32_ts(tb, a_hi + k * 8, bd, idesc_tf32(128, 64), 1u);
                    umma_tf32_ts(tb, a_lo + k * 8, bd, idesc_tf32(128, 32), 1u);
                }
                if (CE > 0 && (it % CE) == CE - 1) umma_commit(bar + st);
                if (++st == 6) { st = 0; need += 4u; }
            }
            umma_commit(bar + 7);
            mbar_wait(bar + 7, 0);
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int CE, int POLL, bool FENCE>
static void run(const char* name, long long* d) {
    const int iters = 4000;
    bench<CE, POLL, FENCE><<<148, 128>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-60s %.1f clk per chunk (tensor floor 192)\n", name, (double)h / iters);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    run<0, 0, false>("8 MMAs", d);
    run<1, 0, true>("fence::after + 8 MMAs + commit", d);
    run<1, 1, true>("ld.acquire poll + fence::after + 8 MMAs + commit", d);
    run<1, 2, true>("ld.relaxed poll + fence::after + 8 MMAs + commit", d);
    run<1, 3, true>("ld.volatile poll + fence::after + 8 MMAs + commit", d);
    run<1, 4, true>("mbarrier.try_wait.relaxed + fence::after + 8 MMAs + commit", d);
    run<1, 5, true>("mbarrier.test_wait.relaxed + fence::after + 8 MMAs + commit", d);
    run<1, 6, true>("mbarrier.try_wait (acquire) + fence::after + 8 MMAs + commit", d);
    run<1, 3, false>("ld.volatile poll + 8 MMAs + commit (no fence)", d);
    return 0;
}
