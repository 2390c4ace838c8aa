// Development check for the tcgen05 spiral-convolution path: runs the FMA kernels and the
// tensor-core kernels through the C ABI on the same random problem, compares them with each
// other and (on a sample of rows) with a double-precision host evaluation, and times both.
//
//   nvcc -O2 -std=c++17 -o tools/umma_check tools/umma_check.cu \
//        -L craniofacialsd-vae_b200 -lsdvae_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../craniofacialsd-vae_b200'
//   tools/umma_check fwd  B V S Cin Cout act iters
//   tools/umma_check bwdx B V S Cin Cout gate iters
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../include/sdvae_b200.h"
extern "C" int sdvae_debug_read_prof(long long* host64);

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)
#define ABI(x)                                                          \
    do {                                                                \
        int rc_ = (x);                                                  \
        if (rc_) {                                                      \
            printf("ABI error %d: %s (%s)\n", rc_, sdvae_last_error(), #x); \
            exit(3);                                                    \
        }                                                               \
    } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 32);
}
static inline float frand() { return (float)(rnd() & 0xFFFFFF) / (float)0x1000000 * 2.f - 1.f; }

template <class T>
static T* dev_copy(const std::vector<T>& h) {
    T* d;
    CK(cudaMalloc(&d, h.size() * sizeof(T) + 16));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}


static void print_prof() {
    if (!getenv("SDVAE_DBG") || !(atoi(getenv("SDVAE_DBG")) & 32)) return;
    long long pr[64];
    sdvae_debug_read_prof(pr);
    printf("prof (CTA 0, cycles)  chunks %lld\n  mma      total %lld  wait a_full %lld  wait t_empty %lld\n"
           "  loader0  total %lld  wait raw_empty %lld  wait copies %lld  wait plan %lld\n"
           "  epilogue total %lld  wait t_full %lld\n  split0   total %lld  wait a_empty %lld  wait raw_full %lld  wait tmem st %lld\n",
           pr[3], pr[0], pr[1], pr[2], pr[8], pr[9], pr[10], pr[11], pr[16], pr[17], pr[24], pr[25], pr[26], pr[27]);
}

struct DevPlan { int* cnt; int* src; int* cell; int rcap; };
static DevPlan make_plan(const std::vector<int>& cell_ptr, const std::vector<int>& cell_src, int out_rows, int S) {
    const int L = sdvae_tc_plan_tiles(out_rows);
    const int mx = sdvae_tc_plan_max_rows(cell_ptr.data(), out_rows, S);
    const int rcap = std::max(32, (mx + 31) / 32 * 32);
    std::vector<int> cnt((size_t)L * S), src((size_t)L * S * rcap), cell((size_t)L * S * 128);
    ABI(sdvae_tc_plan_build(cell_ptr.data(), cell_src.data(), out_rows, S, rcap, cnt.data(), src.data(), cell.data()));
    printf("plan: %d tiles/mesh, max staged rows %d (rcap %d)\n", L, mx, rcap);
    DevPlan d{dev_copy(cnt), dev_copy(src), dev_copy(cell), rcap};
    return d;
}

static double elu_d(double v) { return v > 0 ? v : expm1(v); }

int main(int argc, char** argv) {
    if (argc < 9) {
        printf("usage: %s fwd|bwdx B V S Cin Cout flag iters\n", argv[0]);
        return 1;
    }
    const bool bwdx = strcmp(argv[1], "bwdx") == 0;
    const int B = atoi(argv[2]), V = atoi(argv[3]), S = atoi(argv[4]), Cin = atoi(argv[5]), Cout = atoi(argv[6]);
    const int flag = atoi(argv[7]), iters = atoi(argv[8]);
    printf("%s B=%d V=%d S=%d Cin=%d Cout=%d flag=%d\n", argv[1], B, V, S, Cin, Cout, flag);

    // spiral table: column 0 = the vertex itself, the rest nearby random vertices
    std::vector<int> idx((size_t)V * S);
    for (int v = 0; v < V; ++v) {
        idx[(size_t)v * S] = v;
        for (int s = 1; s < S; ++s) {
            int u = v + (int)(rnd() % 401) - 200;
            if (rnd() % 16 == 0) u = (int)(rnd() % V);
            u = std::min(std::max(u, 0), V - 1);
            idx[(size_t)v * S + s] = u;
        }
    }
    std::vector<float> W((size_t)Cout * S * Cin), bias(Cout);
    for (auto& w : W) w = 0.1f * frand();
    for (auto& b : bias) b = 0.05f * frand();
    cudaStream_t st = 0;
    int* d_idx = dev_copy(idx);
    float* d_W = dev_copy(W);
    float* d_bias = dev_copy(bias);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    if (strcmp(argv[1], "bww") == 0) {
        // ---- weight gradient ------------------------------------------------------------------
        std::vector<int> uptr((size_t)V * S + 1);
        for (size_t i = 0; i <= (size_t)V * S; ++i) uptr[i] = (int)i;
        DevPlan pl = make_plan(uptr, idx, V, S);
        if (!sdvae_tc_bwd_w_supported(S, Cin, Cout, pl.rcap)) { printf("shape not supported by the tcgen05 dW path\n"); return 4; }
        std::vector<float> x((size_t)B * V * Cin), gg((size_t)B * V * Cout);
        for (auto& t : x) t = 1.5f * frand();
        for (auto& t : gg) t = frand();
        float* d_x = dev_copy(x);
        float* d_g = dev_copy(gg);
        const size_t nw = (size_t)Cout * S * Cin;
        float *d_w0, *d_w1, *d_b0, *d_b1, *d_ws;
        CK(cudaMalloc(&d_w0, nw * 4)); CK(cudaMalloc(&d_w1, nw * 4));
        CK(cudaMalloc(&d_b0, Cout * 4)); CK(cudaMalloc(&d_b1, Cout * 4));
        CK(cudaMemset(d_w1, 0xFF, nw * 4)); CK(cudaMemset(d_b1, 0xFF, Cout * 4));
        CK(cudaMalloc(&d_ws, sdvae_spiralconv_bwd_w_workspace((long long)B * V, S, Cin, Cout)));
        ABI(sdvae_spiralconv_bwd_w(d_x, d_idx, d_g, d_w0, d_b0, d_ws, B, V, V, S, Cin, Cout, st));
        ABI(sdvae_spiralconv_bwd_w_tc(d_x, pl.cnt, pl.src, pl.rcap, d_g, d_w1, d_b1, d_ws, B, V, V, S, Cin, Cout, st));
        CK(cudaDeviceSynchronize());
        std::vector<float> w0(nw), w1(nw), b0(Cout), b1(Cout);
        CK(cudaMemcpy(w0.data(), d_w0, nw * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(w1.data(), d_w1, nw * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b0.data(), d_b0, Cout * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b1.data(), d_b1, Cout * 4, cudaMemcpyDeviceToHost));
        double maxref = 0, maxdiff = 0, bref = 0, bdiff = 0; size_t worst = 0;
        for (size_t i = 0; i < nw; ++i) {
            maxref = std::max(maxref, (double)fabsf(w0[i]));
            const double d = fabs((double)w0[i] - (double)w1[i]);
            if (!(d <= maxdiff)) { maxdiff = d; worst = i; }
        }
        for (int i = 0; i < Cout; ++i) { bref = std::max(bref, (double)fabsf(b0[i])); bdiff = std::max(bdiff, fabs((double)b0[i] - b1[i])); }
        printf("dW tc vs fma: max|diff| %.3e  max|ref| %.3e  normwise %.3e (worst at %zu: %g vs %g);  db normwise %.3e\n",
               maxdiff, maxref, maxdiff / maxref, worst, w1[worst], w0[worst], bdiff / bref);
        // fp64 check of a few dW entries
        double e_fma = 0, e_tc = 0, mref = 0;
        for (int t = 0; t < 24; ++t) {
            const int n = (int)(rnd() % Cout), k = (int)(rnd() % (S * Cin));
            const int s = k / Cin, c = k % Cin;
            double acc = 0;
            for (int b = 0; b < B; ++b)
                for (int v = 0; v < V; ++v)
                    acc += (double)gg[((size_t)b * V + v) * Cout + n] * x[((size_t)b * V + idx[(size_t)v * S + s]) * Cin + c];
            mref = std::max(mref, fabs(acc));
            e_fma = std::max(e_fma, fabs(acc - w0[(size_t)n * S * Cin + k]));
            e_tc = std::max(e_tc, fabs(acc - w1[(size_t)n * S * Cin + k]));
        }
        printf("vs fp64 (24 entries): fma %.3e  tc %.3e  (relative to max|ref| %.3e)\n", e_fma / mref, e_tc / mref, mref);
        float ms0, ms1;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) ABI(sdvae_spiralconv_bwd_w(d_x, d_idx, d_g, d_w0, d_b0, d_ws, B, V, V, S, Cin, Cout, st));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms0, e0, e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) ABI(sdvae_spiralconv_bwd_w_tc(d_x, pl.cnt, pl.src, pl.rcap, d_g, d_w1, d_b1, d_ws, B, V, V, S, Cin, Cout, st));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms1, e0, e1));
        const double flops = 2.0 * B * V * (double)S * Cin * Cout;
        printf("time/launch: fma %.3f ms (%.1f TFLOP/s)   tc %.3f ms (%.1f TFLOP/s)   speedup %.2fx\n",
               ms0 / iters, flops / (ms0 / iters) * 1e-9, ms1 / iters, flops / (ms1 / iters) * 1e-9, ms0 / ms1);
        const bool ok = maxdiff / maxref < 2e-5 && bdiff / bref < 2e-5;
        printf(ok ? "CHECK OK\n" : "CHECK FAILED\n");
        return ok ? 0 : 5;
    }
    if (!bwdx) {
        std::vector<int> uptr((size_t)V * S + 1);
        for (size_t i = 0; i <= (size_t)V * S; ++i) uptr[i] = (int)i;
        DevPlan pl = make_plan(uptr, idx, V, S);
        if (!sdvae_tc_supported(S, Cin, Cout, pl.rcap)) { printf("shape not supported by the tcgen05 path\n"); return 4; }
        std::vector<float> x((size_t)B * V * Cin);
        for (auto& t : x) t = 1.5f * frand();
        float* d_x = dev_copy(x);
        float *d_y0, *d_y1, *d_img;
        const size_t ny = (size_t)B * V * Cout;
        CK(cudaMalloc(&d_y0, ny * 4)); CK(cudaMalloc(&d_y1, ny * 4));
        CK(cudaMemset(d_y1, 0xFF, ny * 4));
        CK(cudaMalloc(&d_img, sdvae_tc_wimg_floats(S, Cin, Cout) * 4));
        ABI(sdvae_spiralconv_fwd(d_x, d_idx, d_W, d_bias, d_y0, B, V, V, S, Cin, Cout, flag, st));
        ABI(sdvae_tc_pack_weights(d_W, d_img, S, Cin, Cout, 0, st));
        ABI(sdvae_spiralconv_fwd_tc(d_x, pl.cnt, pl.src, pl.cell, pl.rcap, d_img, d_bias, d_y1, B, V, V, S, Cin, Cout, flag, 0, st));
        CK(cudaDeviceSynchronize());
        std::vector<float> y0(ny), y1(ny);
        CK(cudaMemcpy(y0.data(), d_y0, ny * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(y1.data(), d_y1, ny * 4, cudaMemcpyDeviceToHost));
        double maxref = 0, maxdiff = 0; size_t bad = 0, worst = 0;
        for (size_t i = 0; i < ny; ++i) {
            maxref = std::max(maxref, (double)fabsf(y0[i]));
            const double d = fabs((double)y0[i] - (double)y1[i]);
            if (!(d <= maxdiff)) { maxdiff = d; worst = i; }
            if (!(d < 1e-3)) ++bad;
        }
        printf("tc vs fma: max|diff| %.3e  max|ref| %.3e  normwise %.3e  (elements off by >1e-3: %zu, worst at row %zu col %zu: %g vs %g)\n",
               maxdiff, maxref, maxdiff / maxref, bad, worst / Cout, worst % Cout, y1[worst], y0[worst]);
        // fp64 host check on sampled rows
        double e_fma = 0, e_tc = 0, mref = 0;
        for (int t = 0; t < 512; ++t) {
            const size_t m = ((size_t)rnd() * 2654435761ull) % ((size_t)B * V);
            const int b = (int)(m / V), v = (int)(m % V);
            for (int o = 0; o < Cout; ++o) {
                double acc = bias[o];
                for (int s = 0; s < S; ++s) {
                    const float* xr = &x[((size_t)b * V + idx[(size_t)v * S + s]) * Cin];
                    for (int c = 0; c < Cin; ++c) acc += (double)W[(size_t)o * S * Cin + s * Cin + c] * xr[c];
                }
                if (flag) acc = elu_d(acc);
                mref = std::max(mref, fabs(acc));
                e_fma = std::max(e_fma, fabs(acc - y0[m * Cout + o]));
                e_tc = std::max(e_tc, fabs(acc - y1[m * Cout + o]));
            }
        }
        printf("vs fp64 (512 rows): fma %.3e  tc %.3e  (normwise; max|ref| %.3e)\n", e_fma / mref, e_tc / mref, mref);
        float ms0, ms1;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) ABI(sdvae_spiralconv_fwd(d_x, d_idx, d_W, d_bias, d_y0, B, V, V, S, Cin, Cout, flag, st));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms0, e0, e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) ABI(sdvae_spiralconv_fwd_tc(d_x, pl.cnt, pl.src, pl.cell, pl.rcap, d_img, d_bias, d_y1, B, V, V, S, Cin, Cout, flag, 0, st));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms1, e0, e1));
        const double flops = 2.0 * B * V * (double)S * Cin * Cout, bytes = 4.0 * B * V * (double)(Cin + Cout);
        print_prof();
        printf("time/launch: fma %.3f ms (%.1f TFLOP/s, %.0f GB/s alg)   tc %.3f ms (%.1f TFLOP/s, %.0f GB/s alg)   speedup %.2fx\n",
               ms0 / iters, flops / (ms0 / iters) * 1e-9, bytes / (ms0 / iters) * 1e-6,
               ms1 / iters, flops / (ms1 / iters) * 1e-9, bytes / (ms1 / iters) * 1e-6, ms0 / ms1);
        const bool ok = maxdiff / maxref < 2e-5 && bad == 0;
        printf(ok ? "CHECK OK\n" : "CHECK FAILED\n");
        return ok ? 0 : 5;
    }

    // ---- backward to the input ----------------------------------------------------------------
    std::vector<int> cell_ptr((size_t)V * S + 1, 0), cell_src((size_t)V * S);
    for (int r = 0; r < V; ++r)
        for (int s = 0; s < S; ++s) cell_ptr[(size_t)idx[(size_t)r * S + s] * S + s + 1]++;
    for (size_t i = 0; i < (size_t)V * S; ++i) cell_ptr[i + 1] += cell_ptr[i];
    {
        std::vector<int> fill(cell_ptr.begin(), cell_ptr.end() - 1);
        for (int r = 0; r < V; ++r)
            for (int s = 0; s < S; ++s) cell_src[fill[(size_t)idx[(size_t)r * S + s] * S + s]++] = r;
    }
    DevPlan pl = make_plan(cell_ptr, cell_src, V, S);
    if (!sdvae_tc_supported(S, Cout, Cin, pl.rcap)) { printf("shape not supported by the tcgen05 path\n"); return 4; }
    std::vector<float> dpre((size_t)B * V * Cout), gate((size_t)B * V * Cin);
    for (auto& t : dpre) t = frand();
    for (auto& t : gate) t = frand();            // plays the layer output y: y>0 -> 1, else y+1
    float* d_dpre = dev_copy(dpre);
    float* d_gate = dev_copy(gate);
    int* d_cp = dev_copy(cell_ptr);
    int* d_cs = dev_copy(cell_src);
    float *d_wt, *d_img, *d_dx0, *d_dx1;
    const size_t nx = (size_t)B * V * Cin;
    CK(cudaMalloc(&d_wt, (size_t)Cin * S * Cout * 4));
    CK(cudaMalloc(&d_img, sdvae_tc_wimg_floats(S, Cout, Cin) * 4));
    CK(cudaMalloc(&d_dx0, nx * 4)); CK(cudaMalloc(&d_dx1, nx * 4));
    CK(cudaMemset(d_dx1, 0xFF, nx * 4));
    const float* g = flag ? d_gate : nullptr;
    ABI(sdvae_weight_transpose(d_W, d_wt, Cout, Cin, S, st));
    ABI(sdvae_spiralconv_bwd_x(d_dpre, d_cp, d_cs, d_wt, g, d_dx0, B, V, V, S, Cout, Cin, st));
    ABI(sdvae_tc_pack_weights(d_W, d_img, S, Cin, Cout, 1, st));
    ABI(sdvae_spiralconv_bwd_x_tc(d_dpre, pl.cnt, pl.src, pl.cell, pl.rcap, d_img, g, d_dx1, B, V, V, S, Cout, Cin, 0, st));
    CK(cudaDeviceSynchronize());
    std::vector<float> x0(nx), x1(nx);
    CK(cudaMemcpy(x0.data(), d_dx0, nx * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(x1.data(), d_dx1, nx * 4, cudaMemcpyDeviceToHost));
    double maxref = 0, maxdiff = 0; size_t bad = 0, worst = 0;
    for (size_t i = 0; i < nx; ++i) {
        maxref = std::max(maxref, (double)fabsf(x0[i]));
        const double d = fabs((double)x0[i] - (double)x1[i]);
        if (!(d <= maxdiff)) { maxdiff = d; worst = i; }
        if (!(d < 1e-3)) ++bad;
    }
    printf("tc vs fma: max|diff| %.3e  max|ref| %.3e  normwise %.3e  (elements off by >1e-3: %zu, worst at row %zu col %zu: %g vs %g)\n",
           maxdiff, maxref, maxdiff / maxref, bad, worst / Cin, worst % Cin, x1[worst], x0[worst]);
    double e_fma = 0, e_tc = 0, mref = 0;
    for (int t = 0; t < 512; ++t) {
        const size_t m = ((size_t)rnd() * 2654435761ull) % ((size_t)B * V);
        const int b = (int)(m / V), u = (int)(m % V);
        for (int c = 0; c < Cin; ++c) {
            double acc = 0;
            for (int s = 0; s < S; ++s)
                for (int e = cell_ptr[(size_t)u * S + s]; e < cell_ptr[(size_t)u * S + s + 1]; ++e) {
                    const float* dr = &dpre[((size_t)b * V + cell_src[e]) * Cout];
                    for (int o = 0; o < Cout; ++o) acc += (double)dr[o] * W[(size_t)o * S * Cin + s * Cin + c];
                }
            if (flag) { const float y = gate[m * Cin + c]; acc *= (y > 0.f ? 1.0 : (double)y + 1.0); }
            mref = std::max(mref, fabs(acc));
            e_fma = std::max(e_fma, fabs(acc - x0[m * Cin + c]));
            e_tc = std::max(e_tc, fabs(acc - x1[m * Cin + c]));
        }
    }
    printf("vs fp64 (512 rows): fma %.3e  tc %.3e  (normwise; max|ref| %.3e)\n", e_fma / mref, e_tc / mref, mref);
    float ms0, ms1;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) ABI(sdvae_spiralconv_bwd_x(d_dpre, d_cp, d_cs, d_wt, g, d_dx0, B, V, V, S, Cout, Cin, st));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms0, e0, e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) ABI(sdvae_spiralconv_bwd_x_tc(d_dpre, pl.cnt, pl.src, pl.cell, pl.rcap, d_img, g, d_dx1, B, V, V, S, Cout, Cin, 0, st));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms1, e0, e1));
    const double flops = 2.0 * B * V * (double)S * Cin * Cout;
    print_prof();
    printf("time/launch: fma %.3f ms (%.1f TFLOP/s)   tc %.3f ms (%.1f TFLOP/s)   speedup %.2fx\n",
           ms0 / iters, flops / (ms0 / iters) * 1e-9, ms1 / iters, flops / (ms1 / iters) * 1e-9, ms0 / ms1);
    const bool ok = maxdiff / maxref < 2e-5 && bad == 0;
    printf(ok ? "CHECK OK\n" : "CHECK FAILED\n");
    return ok ? 0 : 5;
}
