// C ABI of sdvae_b200 (see include/sdvae_b200.h).  Host-side dispatch only: argument
// checks, template selection, launches.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include "../../include/sdvae_b200.h"
#include <algorithm>
#include "common.cuh"
#include "spiral_conv.cuh"
#include "spiral_conv_umma.cuh"
#include "spiral_conv_umma_bw.cuh"
#include "spiral_conv_tile.cuh"
#include "spiral_conv_tile_bw.cuh"
#include "spiral_conv_tile_out.cuh"
#include "spiral_conv_tile_out_bw.cuh"
#include "slot_pack.cuh"
#include "pool_misc.cuh"
#include "narrow_conv.cuh"
#include "loss.cuh"

namespace sdvae {
char g_last_error[512] = "";

// true the first time it is asked on each device (cudaFuncSetAttribute is per device, a process may drive several)
struct DeviceOnce {
    unsigned long long mask = 0ull;
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        if (mask & bit) return false;
        mask |= bit;
        return true;
    }
};

// Tuning knobs are read from the environment ONLY in tuning builds (-DSDVAE_TUNING, build.py SDVAE_TUNING=1): the
// shipped library's dispatch does not depend on environment variables.
static int tuning_env(const char* name, int dflt) {
#ifdef SDVAE_TUNING
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
#else
    (void)name;
    return dflt;
#endif
}


static inline unsigned blocks_for(long long n, int threads) {
    return (unsigned)((n + threads - 1) / threads);
}

// ---- tiled gather-contraction dispatch ----------------------------------------
template <int KS, int NT, int TM, int TN, int NWARPS, bool RAGGED>
static int launch_gc(const GcArgs& a, int epi, cudaStream_t st) {
    using Cfg = GcCfg<KS, NT, TM, TN, NWARPS>;
    const size_t smem = Cfg::smem_bytes(a.S, RAGGED);
    const dim3 grid((unsigned)((a.M + Cfg::BM - 1) / Cfg::BM), (unsigned)((a.n_real + NT - 1) / NT));
    auto kern = gc_tile_kernel<KS, NT, TM, TN, NWARPS, RAGGED>;
    static DeviceOnce attr_done;       // the attribute is per device
    if (attr_done.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    GcArgs b = a;
    b.epi = epi;
    kern<<<grid, Cfg::THREADS, smem, st>>>(b);
    return check_launch("gc_tile_kernel");
}

template <bool RAGGED>
static int dispatch_gc(const GcArgs& a, int KS, int epi, cudaStream_t st) {
    if (a.M <= 0) return SDVAE_OK;
    const int N = a.n_real;
    const bool w_ok = (a.ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
    if (w_ok && a.S <= 64) {
        if (KS == 32 && N % 32 == 0 && N % 64 != 0) return launch_gc<32, 32, 4, 8, 8, RAGGED>(a, epi, st);
        if (KS == 32 && N % 64 == 0) return launch_gc<32, 64, 4, 8, 8, RAGGED>(a, epi, st);
        if (KS == 64 && N % 32 == 0 && N % 64 != 0) return launch_gc<64, 32, 4, 8, 8, RAGGED>(a, epi, st);
        if (KS == 64 && N % 64 == 0) return launch_gc<64, 64, 4, 8, 8, RAGGED>(a, epi, st);
        if (KS == 32 && N <= 4) return launch_gc<32, 4, 2, 4, 4, RAGGED>(a, epi, st);
        if (KS == 64 && N <= 4) return launch_gc<64, 4, 2, 4, 4, RAGGED>(a, epi, st);
    }
    if (KS == 3 && a.S * 3 <= 32 && N % 32 == 0) return launch_gc<3, 32, 4, 8, 8, RAGGED>(a, epi, st);
    // shape-agnostic fallback
    const long long total = a.M * a.n_real;
    gc_generic_kernel<RAGGED><<<blocks_for(total, 256), 256, 0, st>>>(a, KS, epi);
    return check_launch("gc_generic_kernel");
}

// ---- weight gradient dispatch ---------------------------------------------------
constexpr int kBwBMW = 16;
constexpr int kBwMaxSplit = 592;     // 4 CTAs per SM
constexpr int kBwMinRows = 256;

static inline void bw_split(long long M, long long* rows_per_cta, int* nsplit) {
    long long rows = (M + kBwMaxSplit - 1) / kBwMaxSplit;
    if (rows < kBwMinRows) rows = kBwMinRows;
    rows = (rows + kBwBMW - 1) / kBwBMW * kBwBMW;
    *rows_per_cta = rows;
    *nsplit = (int)((M + rows - 1) / rows);
    if (*nsplit < 1) *nsplit = 1;
}

template <int KS, int S_, int NT, int TK, int TN>
static int launch_bw(const BwArgs& a, int nsplit, cudaStream_t st) {
    using Cfg = BwCfg<KS, S_, NT, TK, TN, kBwBMW>;
    auto kern = bw_outer_kernel<KS, S_, NT, TK, TN, kBwBMW>;
    static DeviceOnce attr_done;       // the attribute is per device
    if (attr_done.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    kern<<<nsplit, Cfg::THREADS, Cfg::SMEM, st>>>(a);
    return check_launch("bw_outer_kernel");
}

// ---- tcgen05 (tensor-core) gather-contraction dispatch ------------------------------------
static inline int tc_tile_n(int N) { return N <= 16 ? 16 : (N <= 32 ? 32 : (N <= 64 ? 64 : 0)); }

// raw ring depth for a layer shape (0 = the resident weight image leaves no room)
static int tc_raw_stages(int S, int KS, int N, int rcap) {
    const int NT = tc_tile_n(N);
    if (NT == 0) return 0;
    const long long b_bytes = (long long)S * (KS / 32) * 2 * NT * 128;
    const long long budget = 226LL * 1024 - 2048 - b_bytes;
    long long st = budget / ((long long)rcap * 128);
    if (st > umma::kMaxRaw) st = umma::kMaxRaw;
    return st < 0 ? 0 : (int)st;
}

static bool tc_shape_ok(int S, int KS, int N, int rcap) {
    if (KS != 32 && KS != 64) return false;
    if (N < 1 || tc_tile_n(N) == 0 || S < 1) return false;
    if (rcap < 32 || rcap > umma::kMaxRcap || rcap % 32 != 0) return false;
    return tc_raw_stages(S, KS, N, rcap) >= 3;      // NS <= NAST <= NRAW: fewer stages = fewer active splitter sets
}

template <int KS, int NT, bool UNIFORM>
static int launch_umma(umma::UmmaArgs& ua, cudaStream_t st) {
    using Cfg = umma::UmmaCfg<KS, NT>;
    auto kern = umma::gc_umma_kernel<KS, NT, UNIFORM>;
    static DeviceOnce attr_done;       // the attribute is per device
    if (attr_done.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    ua.nraw = Cfg::raw_stages(ua.S, ua.rcap);
    { static int dbg = -1; if (dbg < 0) dbg = tuning_env("SDVAE_DBG", 0); ua.dbg = dbg; }
    const long long ntiles = (long long)ua.B * ua.L;
    const int grid = ntiles < kNumSMs ? (int)ntiles : kNumSMs;
    ua.ostage = (UNIFORM && Cfg::out_stage(ua.S, ua.rcap)) ? 1 : 0;
    kern<<<grid, umma::kThreads, Cfg::smem_bytes(ua.S, ua.rcap, ua.nraw, ua.ostage != 0), st>>>(ua);
    return check_launch("gc_umma_kernel");
}

template <bool UNIFORM>
static int dispatch_umma(umma::UmmaArgs& ua, int KS, cudaStream_t st) {
    if (ua.B <= 0) return SDVAE_OK;
    const int NT = tc_tile_n(ua.n_real);
    if (KS == 32 && NT == 16) return launch_umma<32, 16, UNIFORM>(ua, st);
    if (KS == 32 && NT == 32) return launch_umma<32, 32, UNIFORM>(ua, st);
    if (KS == 32 && NT == 64) return launch_umma<32, 64, UNIFORM>(ua, st);
    if (KS == 64 && NT == 16) return launch_umma<64, 16, UNIFORM>(ua, st);
    if (KS == 64 && NT == 32) return launch_umma<64, 32, UNIFORM>(ua, st);
    if (KS == 64 && NT == 64) return launch_umma<64, 64, UNIFORM>(ua, st);
    return set_error(SDVAE_ERR_UNSUPPORTED, "tcgen05 path: unsupported layer shape");
}

template <bool RAGGED, int NSETS>
static int launch_tile_n(tile::TileArgs& ta, cudaStream_t st) {
    auto kern = tile::gt_kernel<RAGGED, NSETS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // per device, cheap
    ta.nts = tile::TileCfg::stages(ta.S, ta.rcap, ta.ecap);
    { static int dbg = -1; if (dbg < 0) dbg = tuning_env("SDVAE_DBG", 0); ta.dbg = dbg; }
    const long long ntiles = (long long)ta.B * ta.L;
    const int grid = ntiles < kNumSMs ? (int)ntiles : kNumSMs;
    kern<<<grid, tile::TileWarps<NSETS>::kThreads, tile::TileCfg::smem_bytes(ta.S, ta.rcap, ta.ecap, ta.nts), st>>>(ta);
    return check_launch("gt_kernel");
}

template <bool RAGGED>
static int launch_tile(tile::TileArgs& ta, cudaStream_t st) {
    // four splitter sets: five and six (fewer registers per thread) measured slower (profiles/r02_tile_kernel.md)
    return launch_tile_n<RAGGED, 4>(ta, st);
}

// Meshes per CTA (MG) of the staged kernels: runs long enough to amortise the ring's fill, and a CTA count that
// fills whole waves of the resident CTAs (a wave lasts ~MG mesh-tiles: minimise waves x (MG + fill)).
static int pick_mesh_group(int B, int L, long long slots, int nst) {
    int MG = 1;
    long long best = -1;
    for (int mg = std::min(B, 2 * nst); mg <= std::min(B, 48); ++mg) {
        const long long ctas = (long long)L * ((B + mg - 1) / mg);
        const long long cost = ((ctas + slots - 1) / slots) * (mg + nst);
        if (best < 0 || cost < best) { best = cost; MG = mg; }
    }
    return MG;
}

template <int CQ, int WD, int NST>
static int launch_pool_staged_n(const float* x, const int32_t* tile_ptr, const int32_t* stage_src,
                                const int32_t* ent, float* out, int B, int Vin, int Vout, int ucap,
                                cudaStream_t st) {
    auto kern = pool_ell_fwd_staged_kernel<CQ, WD, NST>;
    static DeviceOnce attr_done;       // the attribute is per device
    if (attr_done.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const size_t smem = (size_t)NST * ucap * CQ * 16;
    const int L = (Vout + kPoolTile - 1) / kPoolTile;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (226 * 1024) / (smem + 1024)));
    const int MG = pick_mesh_group(B, L, (long long)kNumSMs * per_sm, NST);
    const long long grid = (long long)L * ((B + MG - 1) / MG);
    kern<<<(unsigned)grid, kPoolStageThreads, smem, st>>>(x, tile_ptr, stage_src,
                                                          reinterpret_cast<const int2*>(ent), out, B, Vin,
                                                          Vout, L, ucap, MG);
    return check_launch("pool_ell_fwd_staged_kernel");
}

// ring depth: 128-byte rows (C = 32) run 3 CTAs per SM with 3 stages (measured 0.456 ms against 0.468 ms with
// 4 stages / 2 CTAs at 1024 meshes, level 0); 256-byte rows (C = 64, 2 CTAs per SM either way) want the deeper ring
template <int CQ, int WD>
static int launch_pool_staged(const float* x, const int32_t* tile_ptr, const int32_t* stage_src,
                              const int32_t* ent, float* out, int B, int Vin, int Vout, int ucap,
                              cudaStream_t st) {
    return launch_pool_staged_n<CQ, WD, pool_stages_for(CQ)>(x, tile_ptr, stage_src, ent, out, B, Vin, Vout, ucap, st);
}

template <int CQ>
static int launch_pool_staged_w(const float* x, const int32_t* tile_ptr, const int32_t* stage_src,
                                const int32_t* ent, float* out, int B, int Vin, int Vout, int Wd, int ucap,
                                cudaStream_t st) {
    switch (Wd) {
        case 2: return launch_pool_staged<CQ, 2>(x, tile_ptr, stage_src, ent, out, B, Vin, Vout, ucap, st);
        case 3: return launch_pool_staged<CQ, 3>(x, tile_ptr, stage_src, ent, out, B, Vin, Vout, ucap, st);
        default: return launch_pool_staged<CQ, 4>(x, tile_ptr, stage_src, ent, out, B, Vin, Vout, ucap, st);
    }
}
}  // namespace sdvae

using namespace sdvae;

extern "C" {

int sdvae_abi_version(void) { return 1; }
const char* sdvae_last_error(void) { return g_last_error; }

int sdvae_spiralconv_fwd(const float* x, const int32_t* idx, const float* W, const float* bias,
                         float* y, int B, int Vin, int Vout, int S, int Cin, int Cout, int act,
                         sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && idx && W && y, "spiralconv_fwd: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0 && S > 0 && Cin > 0 && Cout > 0, "spiralconv_fwd: bad shape");
    SDVAE_REQUIRE((long long)B * Vin < 2147483647LL, "spiralconv_fwd: B*Vin exceeds int32 rows");
    GcArgs a{};
    a.in = x; a.idx = idx; a.W = W; a.bias = bias; a.out = y;
    a.M = (long long)B * Vout; a.in_rows = Vin; a.Vout = Vout; a.S = S;
    a.ldw = S * Cin; a.ldo = Cout; a.n_real = Cout;
    const int epi = act == SDVAE_ACT_ELU ? EPI_BIAS_ELU : EPI_BIAS;
    return dispatch_gc<false>(a, Cin, epi, (cudaStream_t)stream);
}

int sdvae_weight_transpose(const float* W, float* Wt, int Cout, int Cin, int S, sdvae_stream_t stream) {
    SDVAE_REQUIRE(W && Wt && Cout > 0 && Cin > 0 && S > 0, "weight_transpose: bad argument");
    const int total = Cout * Cin * S;
    weight_transpose_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(W, Wt, Cout, Cin, S);
    return check_launch("weight_transpose_kernel");
}

int sdvae_spiralconv_bwd_x(const float* dpre, const int32_t* cell_ptr, const int32_t* cell_src,
                           const float* Wt, const float* gate, float* dx, int B, int Vrows,
                           int Vdst, int S, int Cout, int Cin, sdvae_stream_t stream) {
    SDVAE_REQUIRE(dpre && cell_ptr && cell_src && Wt && dx, "spiralconv_bwd_x: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vrows > 0 && Vdst > 0 && S > 0 && Cin > 0 && Cout > 0, "spiralconv_bwd_x: bad shape");
    SDVAE_REQUIRE((long long)B * Vrows < 2147483647LL, "spiralconv_bwd_x: B*Vrows exceeds int32 rows");
    GcArgs a{};
    a.in = dpre; a.cell_ptr = cell_ptr; a.cell_src = cell_src; a.W = Wt; a.gate = gate; a.out = dx;
    a.M = (long long)B * Vdst; a.in_rows = Vrows; a.Vout = Vdst; a.S = S;
    a.ldw = S * Cout; a.ldo = Cin; a.n_real = Cin;
    return dispatch_gc<true>(a, Cout, gate ? EPI_GATE : EPI_NONE, (cudaStream_t)stream);
}

/* tuning aid (not part of the product interface): cycle counters written by CTA 0 when SDVAE_DBG has bit 32 */
int sdvae_debug_read_prof(long long* host64) {
    return cudaMemcpyFromSymbol(host64, umma::g_prof, 64 * sizeof(long long)) == cudaSuccess ? 0 : 2;
}

int sdvae_tc_supported(int S, int KS, int N, int rcap) { return tc_shape_ok(S, KS, N, rcap) ? 1 : 0; }

size_t sdvae_tc_wimg_floats(int S, int KS, int N) {
    const int NT = tc_tile_n(N);
    return (size_t)S * (KS / 32) * 2 * NT * 32;
}

int sdvae_tc_pack_weights(const float* W, float* wimg, int S, int Cin, int Cout, int transposed,
                          sdvae_stream_t stream) {
    return sdvae_tc_pack_weights_part(W, wimg, S, Cin, Cout, transposed, 0, (transposed & 1) ? Cin : Cout, stream);
}

int sdvae_tc_pack_weights_part(const float* W, float* wimg, int S, int Cin, int Cout, int transposed,
                               int n0, int n_cnt, sdvae_stream_t stream) {
    SDVAE_REQUIRE(W && wimg && S > 0 && Cin > 0 && Cout > 0, "tc_pack_weights: bad argument");
    const int tr = transposed & 1;       // bit 1 of the flag word: kperm image (tile-staged kernels)
    const int KS = tr ? Cout : Cin, Nfull = tr ? Cin : Cout, N = n_cnt;
    SDVAE_REQUIRE(n0 >= 0 && n_cnt > 0 && n0 + n_cnt <= Nfull, "tc_pack_weights: bad output-channel range");
    if (!tc_shape_ok(S, KS, N, 128)) return set_error(SDVAE_ERR_UNSUPPORTED, "tc_pack_weights: unsupported layer shape");
    SDVAE_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0, "tc_pack_weights: wimg must be 16-byte aligned");
    umma::PackArgs a;
    // rows n0 .. n0+n_cnt of the (transposed) weight: forward W[n, k] -> offset n0 rows; transposed
    // Wt[c, s*Cout + o] = W[o, s*Cin + c] -> offset n0 columns
    a.W = tr ? W + n0 : W + (size_t)n0 * S * Cin;
    a.img = wimg; a.NT = tc_tile_n(N); a.KS = KS; a.S = S; a.n_real = N; a.ldw = S * Cin;
    a.transposed = transposed & 3; a.cin = Cin;
    const long long total = (long long)S * (KS / 32) * 2 * a.NT * 32;
    umma::umma_pack_weights_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("umma_pack_weights_kernel");
}

int sdvae_tc_pack_weights_batch(const sdvae_pack_entry* entries, int n, sdvae_stream_t stream) {
    SDVAE_REQUIRE(entries && n >= 0, "tc_pack_weights_batch: bad argument");
    static_assert(sizeof(sdvae_pack_entry) == sizeof(umma::PackEntry), "pack entry layouts differ");
    if (n == 0) return SDVAE_OK;
    // the caller validated every entry when it built the table (sdvae_tc_supported); 32 blocks x 256 threads
    // cover the largest image (64 x 576 weights -> 2 x 64 x 576 floats) in a few strides
    umma::umma_pack_weights_batch_kernel<<<dim3(32, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const umma::PackEntry*>(entries));
    return check_launch("umma_pack_weights_batch_kernel");
}

int sdvae_tc_plan_tiles(int out_rows) { return (out_rows + umma::kBM - 1) / umma::kBM; }

int sdvae_tc_plan_max_rows(const int32_t* cell_ptr, int out_rows, int S) {
    if (!cell_ptr || out_rows <= 0 || S <= 0) return -1;
    const int L = sdvae_tc_plan_tiles(out_rows);
    int mx = 0;
    for (int jt = 0; jt < L; ++jt)
        for (int s = 0; s < S; ++s) {
            int n = 0;
            const int r1 = (jt + 1) * umma::kBM < out_rows ? (jt + 1) * umma::kBM : out_rows;
            for (int r = jt * umma::kBM; r < r1; ++r) n += cell_ptr[(size_t)r * S + s + 1] - cell_ptr[(size_t)r * S + s];
            if (n > mx) mx = n;
        }
    return mx;
}

int sdvae_tc_plan_build(const int32_t* cell_ptr, const int32_t* cell_src, int out_rows, int S,
                        int rcap, int32_t* cnt, int32_t* src, int32_t* cell) {
    SDVAE_REQUIRE(cell_ptr && cell_src && cnt && src && cell, "tc_plan_build: null pointer");
    SDVAE_REQUIRE(out_rows > 0 && S > 0 && rcap >= 32 && rcap % 32 == 0 && rcap <= umma::kMaxRcap, "tc_plan_build: bad shape");
    const int L = sdvae_tc_plan_tiles(out_rows);
    int32_t rows[umma::kMaxRcap];
    for (int jt = 0; jt < L; ++jt)
        for (int s = 0; s < S; ++s) {
            int n = 0;
            int32_t* crow = cell + ((size_t)jt * S + s) * umma::kBM;
            for (int lr = 0; lr < umma::kBM; ++lr) {
                const int r = jt * umma::kBM + lr;
                int c = 0;
                if (r < out_rows) {
                    const int e0 = cell_ptr[(size_t)r * S + s], e1 = cell_ptr[(size_t)r * S + s + 1];
                    c = e1 - e0;
                    SDVAE_REQUIRE(c >= 0 && n + c <= rcap, "tc_plan_build: a (tile, slot) stages more rows than rcap");
                    for (int e = e0; e < e1; ++e) {
                        SDVAE_REQUIRE(cell_src[e] >= 0 && cell_src[e] < 65536, "tc_plan_build: source row does not fit 16 bits");
                        rows[n + (e - e0)] = cell_src[e];
                    }
                }
                crow[lr] = (int32_t)((uint32_t)n | ((uint32_t)c << 16));
                n += c;
            }
            for (int e = n; e < rcap; ++e) rows[e] = 0;
            cnt[(size_t)jt * S + s] = n;
            // pack in loader-lane order (umma::plan_fetch): staged row e = 32*j + 4*t + rsub -> word
            // 16*j + 4*rsub + (t >> 1), low half for even t
            uint32_t* w = reinterpret_cast<uint32_t*>(src) + ((size_t)jt * S + s) * (rcap / 2);
            for (int j = 0; 32 * j < rcap; ++j)
                for (int rsub = 0; rsub < 4; ++rsub)
                    for (int t = 0; t < 8; t += 2) {
                        const uint32_t lo = (uint32_t)rows[32 * j + 4 * t + rsub];
                        const uint32_t hi = (uint32_t)rows[32 * j + 4 * (t + 1) + rsub];
                        w[16 * j + 4 * rsub + (t >> 1)] = lo | (hi << 16);
                    }
        }
    return SDVAE_OK;
}

static int tc_conv(const float* in, const int32_t* plan_cnt, const int32_t* plan_src,
                   const int32_t* plan_cell, int rcap, const float* wimg, const float* bias,
                   const float* gate, float* out, int B, int in_rows, int out_rows, int S, int KS,
                   int N, int ldo, int epi, bool uniform, cudaStream_t st, const char* who) {
    SDVAE_REQUIRE(in && plan_cnt && plan_src && (uniform || plan_cell) && wimg && out, "spiralconv tc: null pointer");
    SDVAE_REQUIRE(B >= 0 && in_rows > 0 && out_rows > 0 && S > 0 && KS > 0 && N > 0, "spiralconv tc: bad shape");
    SDVAE_REQUIRE((long long)B * in_rows < 2147483647LL, "spiralconv tc: B*rows exceeds int32");
    SDVAE_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(wimg)) & 15) == 0,
                  "spiralconv tc: input and wimg must be 16-byte aligned");
    if (!tc_shape_ok(S, KS, N, rcap)) return set_error(SDVAE_ERR_UNSUPPORTED, who);
    if (tc_tile_n(N) >= 32) {       // full-width tiles use 16-byte epilogue accesses
        const uintptr_t al = reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
                             reinterpret_cast<uintptr_t>(gate);
        if (N != tc_tile_n(N) || (al & 15) != 0 || (ldo > 0 && ldo % 4 != 0)) return set_error(SDVAE_ERR_UNSUPPORTED, who);
    }
    umma::UmmaArgs ua{};
    ua.in = in; ua.plan_cnt = plan_cnt; ua.plan_src = plan_src; ua.plan_cell = plan_cell;
    ua.wimg = wimg; ua.bias = bias; ua.gate = gate; ua.out = out;
    ua.B = B; ua.in_rows = in_rows; ua.out_rows = out_rows; ua.L = sdvae_tc_plan_tiles(out_rows);
    ua.S = S; ua.rcap = rcap; ua.n_real = N; ua.ldo = ldo > 0 ? ldo : N; ua.epi = epi;
    return uniform ? dispatch_umma<true>(ua, KS, st) : dispatch_umma<false>(ua, KS, st);
}

int sdvae_spiralconv_fwd_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                            const int32_t* plan_cell, int rcap, const float* wimg, const float* bias,
                            float* y, int B, int Vin, int Vout, int S, int Cin, int Cout, int act,
                            int ldy, sdvae_stream_t stream) {
    const int epi = act == SDVAE_ACT_ELU ? EPI_BIAS_ELU : EPI_BIAS;
    return tc_conv(x, plan_cnt, plan_src, plan_cell, rcap, wimg, bias, nullptr, y, B, Vin, Vout, S, Cin,
                   Cout, ldy, epi, true, (cudaStream_t)stream, "spiralconv_fwd_tc: unsupported layer shape");
}

int sdvae_spiralconv_bwd_x_tc(const float* dpre, const int32_t* plan_cnt, const int32_t* plan_src,
                              const int32_t* plan_cell, int rcap, const float* wimg_t, const float* gate,
                              float* dx, int B, int Vrows, int Vdst, int S, int Cout, int Cin,
                              int lddx, sdvae_stream_t stream) {
    return tc_conv(dpre, plan_cnt, plan_src, plan_cell, rcap, wimg_t, nullptr, gate, dx, B, Vrows, Vdst, S,
                   Cout, Cin, lddx, gate ? EPI_GATE : EPI_NONE, false, (cudaStream_t)stream,
                   "spiralconv_bwd_x_tc: unsupported layer shape");
}

/* ---- tcgen05 SpiralConv with tile-local staging and stage-granular hand-offs (spiral_conv_tile.cuh) -------- */
int sdvae_tile_supported(int S, int Cin, int Cout, int rcap, int ecap) {
    if (Cin != 32 || Cout != 32 || S < 6 || S > 42 || S % tile::kTChunksPerStage != 0) return 0;   // every splitter set needs a slot in every tile
    if (rcap < 32 || rcap > tile::kTMaxRcap || rcap % 32 != 0) return 0;
    if (ecap < 0 || ecap % 64 != 0 || ecap > 1984) return 0;     // 128-byte aligned tile stages; 11-bit offsets into plan_ext
    return tile::TileCfg::stages(S, rcap, ecap) >= 2 ? 1 : 0;
}

static int tile_conv(const float* in, const int32_t* plan_cnt, const int32_t* plan_src, const uint32_t* plan_cell,
                     const uint16_t* plan_ext, int rcap, int ecap, const float* wimg, const float* bias,
                     const float* gate, float* out, int B, int in_rows, int out_rows, int S, int epi, bool ragged,
                     cudaStream_t st, const char* who) {
    SDVAE_REQUIRE(in && plan_cnt && plan_src && plan_cell && wimg && out && (!ragged || plan_ext || ecap == 0),
                  "spiralconv tile: null pointer");
    SDVAE_REQUIRE(B >= 0 && in_rows > 0 && out_rows > 0, "spiralconv tile: bad shape");
    SDVAE_REQUIRE((long long)B * in_rows < 2147483647LL && (long long)B * out_rows < 2147483647LL,
                  "spiralconv tile: B*rows exceeds int32");
    const uintptr_t al = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(wimg) |
                         reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
                         reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(plan_src) |
                         reinterpret_cast<uintptr_t>(plan_cell) | reinterpret_cast<uintptr_t>(plan_ext);
    SDVAE_REQUIRE((al & 15) == 0, "spiralconv tile: tensors and plan tables must be 16-byte aligned");
    if (!sdvae_tile_supported(S, 32, 32, rcap, ecap)) return set_error(SDVAE_ERR_UNSUPPORTED, who);
    if (B == 0) return SDVAE_OK;
    tile::TileArgs ta{};
    ta.in = in; ta.plan_cnt = plan_cnt; ta.plan_src = plan_src; ta.plan_cell = plan_cell; ta.plan_ext = plan_ext;
    ta.wimg = wimg; ta.bias = bias; ta.gate = gate; ta.out = out;
    ta.B = B; ta.in_rows = in_rows; ta.out_rows = out_rows; ta.L = sdvae_tc_plan_tiles(out_rows);
    ta.S = S; ta.rcap = rcap; ta.ecap = ecap; ta.ldo = 32; ta.epi = epi;
    return ragged ? launch_tile<true>(ta, st) : launch_tile<false>(ta, st);
}

int sdvae_spiralconv_fwd_tile(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                              const uint32_t* plan_cell, int rcap, const float* wimg, const float* bias, float* y,
                              int B, int Vin, int Vout, int S, int Cin, int Cout, int act, sdvae_stream_t stream) {
    if (Cin != 32 || Cout != 32) return set_error(SDVAE_ERR_UNSUPPORTED, "spiralconv_fwd_tile: unsupported layer shape");
    const int epi = act == SDVAE_ACT_ELU ? EPI_BIAS_ELU : EPI_BIAS;
    return tile_conv(x, plan_cnt, plan_src, plan_cell, nullptr, rcap, 0, wimg, bias, nullptr, y, B, Vin, Vout, S, epi,
                     false, (cudaStream_t)stream, "spiralconv_fwd_tile: unsupported layer shape");
}

int sdvae_spiralconv_bwd_x_tile(const float* dpre, const int32_t* plan_cnt, const int32_t* plan_src,
                                const uint32_t* plan_cell, const uint16_t* plan_ext, int rcap, int ecap,
                                const float* wimg_t, const float* gate, float* dx, int B, int Vrows, int Vdst, int S,
                                int Cout, int Cin, sdvae_stream_t stream) {
    if (Cin != 32 || Cout != 32) return set_error(SDVAE_ERR_UNSUPPORTED, "spiralconv_bwd_x_tile: unsupported layer shape");
    return tile_conv(dpre, plan_cnt, plan_src, plan_cell, plan_ext, rcap, ecap, wimg_t, nullptr, gate, dx, B, Vrows, Vdst,
                     S, gate ? EPI_GATE : EPI_NONE, true, (cudaStream_t)stream,
                     "spiralconv_bwd_x_tile: unsupported layer shape");
}

/* ---- narrow-output layer forward on tcgen05, project-then-gather (spiral_conv_tile_out.cuh) ------------------- */
int sdvae_narrow_out_fwd_tc_supported(int S, int Cin, int Cout, int rcap) {
    if (Cin != 32 || Cout < 1 || Cout > 4 || S < 1 || S > 9 || S * Cout > 27) return 0;   // 29 floats per projected row, 27 used + 2 zero
    if (rcap < 32 || rcap > tile::kOMaxRcap || rcap % 32 != 0) return 0;
    return tile::OutCfg::stages(S, rcap) >= 2 ? 1 : 0;
}

int sdvae_narrow_out_fwd_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                            const uint32_t* plan_cell, int rcap, const float* W, const float* bias, float* y,
                            int B, int Vin, int Vout, int S, int Cin, int Cout, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && plan_cnt && plan_src && plan_cell && W && y, "narrow_out_fwd_tc: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0, "narrow_out_fwd_tc: bad shape");
    SDVAE_REQUIRE((long long)B * Vin < 2147483647LL && (long long)B * Vout < 2147483647LL,
                  "narrow_out_fwd_tc: B*rows exceeds int32");
    SDVAE_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(plan_src) |
                    reinterpret_cast<uintptr_t>(plan_cell)) & 15) == 0,
                  "narrow_out_fwd_tc: x and the plan tables must be 16-byte aligned");
    if (!sdvae_narrow_out_fwd_tc_supported(S, Cin, Cout, rcap))
        return set_error(SDVAE_ERR_UNSUPPORTED, "narrow_out_fwd_tc: unsupported layer shape");
    if (B == 0) return SDVAE_OK;
    tile::OutArgs a{};
    a.in = x; a.plan_cnt = plan_cnt; a.plan_src = plan_src; a.plan_cell = plan_cell; a.W = W; a.bias = bias; a.out = y;
    a.B = B; a.in_rows = Vin; a.out_rows = Vout; a.L = sdvae_tc_plan_tiles(Vout); a.S = S; a.NO = Cout; a.rcap = rcap;
    a.nts = tile::OutCfg::stages(S, rcap);
    const bool fixed = S == 9 && Cout == 3;
    auto kern = fixed ? tile::pt_kernel<9, 3> : tile::pt_kernel<0, 0>;
    static DeviceOnce attr_done[2];
    if (attr_done[fixed].first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const long long ntiles = (long long)B * a.L;
    const int grid = ntiles < kNumSMs ? (int)ntiles : kNumSMs;
    kern<<<grid, tile::kOThreads, tile::OutCfg::smem_bytes(S, rcap, a.nts), (cudaStream_t)stream>>>(a);
    return check_launch("narrow_out_fwd_tc");
}

/* ---- narrow-output layer backward on tcgen05, gather-then-project (spiral_conv_tile_out_bw.cuh) --------------- */
int sdvae_narrow_out_bwd_tc_supported(int S, int Cin, int Cout, int rcap, int ecap) {
    if (Cin != 32 || Cout < 1 || Cout > 3 || S < 1 || S > 9 || S * Cout > 32) return 0;
    if (rcap < 32 || rcap > tile::kTMaxRcap || rcap % 32 != 0) return 0;
    if (ecap < 0 || ecap % 64 != 0 || ecap > 1984) return 0;
    return tile::OutBwCfg::stages(S, rcap, ecap) >= 2 ? 1 : 0;
}

size_t sdvae_narrow_out_bwd_tc_workspace(int S, int Cout) {
    return sizeof(float) * (size_t)kNumSMs * ((size_t)Cout * S * 32 + Cout);
}

int sdvae_narrow_out_bwd_tc(const float* dy, const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                            const uint32_t* plan_cell, const uint16_t* plan_ext, int rcap, int ecap, const float* W,
                            float* dx, float* dW, float* db, void* workspace, int B, int Vrows, int Vdst, int S,
                            int Cin, int Cout, int gate, sdvae_stream_t stream) {
    SDVAE_REQUIRE(dy && x && plan_cnt && plan_src && plan_cell && (plan_ext || ecap == 0) && W && dx && dW && db && workspace,
                  "narrow_out_bwd_tc: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vrows > 0 && Vdst > 0, "narrow_out_bwd_tc: bad shape");
    SDVAE_REQUIRE((long long)B * Vrows < 2147483647LL && (long long)B * Vdst < 2147483647LL,
                  "narrow_out_bwd_tc: B*rows exceeds int32");
    SDVAE_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(plan_src) |
                    reinterpret_cast<uintptr_t>(plan_cell) | reinterpret_cast<uintptr_t>(plan_ext)) & 15) == 0,
                  "narrow_out_bwd_tc: x, dx and the plan tables must be 16-byte aligned");
    if (!sdvae_narrow_out_bwd_tc_supported(S, Cin, Cout, rcap, ecap))
        return set_error(SDVAE_ERR_UNSUPPORTED, "narrow_out_bwd_tc: unsupported layer shape");
    cudaStream_t st = (cudaStream_t)stream;
    const long long len = (long long)Cout * S * 32;
    if (B == 0) {
        cudaMemsetAsync(dW, 0, sizeof(float) * len, st);
        cudaMemsetAsync(db, 0, sizeof(float) * Cout, st);
        return check_launch("narrow_out_bwd_tc memset");
    }
    tile::OutBwArgs a{};
    a.dy = dy; a.x = x; a.plan_cnt = plan_cnt; a.plan_src = plan_src; a.plan_cell = plan_cell; a.plan_ext = plan_ext;
    a.W = W; a.dx = dx;
    a.B = B; a.rows_v = Vrows; a.rows_u = Vdst; a.L = sdvae_tc_plan_tiles(Vdst); a.S = S; a.NO = Cout;
    a.rcap = rcap; a.ecap = ecap; a.nts = tile::OutBwCfg::stages(S, rcap, ecap); a.flush = 2; a.gate = gate;
    const long long ntiles = (long long)B * a.L;
    const int grid = ntiles < kNumSMs ? (int)ntiles : kNumSMs;
    a.part = static_cast<float*>(workspace);
    a.part_b = a.part + (size_t)grid * len;
    const bool fixed = S == 9 && Cout == 3;
    auto kern = fixed ? tile::qt_kernel<9, 3> : tile::qt_kernel<0, 0>;
    static DeviceOnce attr_done[2];
    if (attr_done[fixed].first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    kern<<<grid, tile::kQThreads, tile::OutBwCfg::smem_bytes(S, rcap, ecap, a.nts), st>>>(a);
    split_reduce2_kernel<<<blocks_for(len + Cout, 64), 64 * kSplitGroups, 0, st>>>(a.part, a.part_b, dW, db, grid, len, Cout);
    return check_launch("narrow_out_bwd_tc");
}

int sdvae_dense_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src, int rcap,
                   const float* wimg, const float* bias, const float* gate, float* y, int B, int R,
                   int act, sdvae_stream_t stream) {
    const int epi = gate ? EPI_GATE : (act == SDVAE_ACT_ELU ? EPI_BIAS_ELU : EPI_BIAS);
    return tc_conv(x, plan_cnt, plan_src, nullptr, rcap, wimg, gate ? nullptr : bias, gate, y, B, R, R, 1, 32,
                   32, 0, epi, true, (cudaStream_t)stream, "dense_tc: unsupported shape");
}

int sdvae_slot_pack(const float* in, const int32_t* cell_ptr, const int32_t* cell_src, float* out, int B,
                    int Vin, int R, int S, int C, sdvae_stream_t stream) {
    SDVAE_REQUIRE(in && cell_src && out, "slot_pack: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && R > 0 && S > 0 && C > 0 && S * C <= 32, "slot_pack: bad shape (S*C must be <= 32)");
    if (B == 0) return SDVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t in_bytes = (size_t)Vin * C * sizeof(float);
    if (in_bytes <= 220 * 1024) {            // the mesh's narrow input fits in shared memory
        static DeviceOnce attr_done;       // the attribute is per device
        if (attr_done.first()) cudaFuncSetAttribute(slot_pack_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        int parts = 1;                         // split meshes into row ranges until the grid fills the SMs evenly
        while ((long long)B * parts < 4LL * kNumSMs && parts < 16 && R / (parts * 2) >= 256) parts *= 2;
        const long long items = (long long)B * parts;
        const int grid = items < kNumSMs ? (int)items : kNumSMs;
        slot_pack_smem_kernel<<<grid, 1024, in_bytes + 16, st>>>(in, cell_ptr, cell_src, out, B, parts, R, Vin, S, C);
        return check_launch("slot_pack_smem_kernel");
    }
    const long long rows = (long long)B * R;
    // one warp per row, no grid-stride cap: a row is three dependent loads (cell range, source row, value),
    // so the kernel lives on the number of rows in flight
    long long blocks = (rows + 8 * kSlotRows - 1) / (8 * kSlotRows);      // 8 warps x kSlotRows rows per block
    if (blocks > 0x7fffffffLL) blocks = 0x7fffffffLL;
    slot_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, cell_ptr, cell_src, out, rows, R, Vin, S, C);
    return check_launch("slot_pack_kernel");
}

int sdvae_slot_weight(const float* W, float* Wd, int mode, int N, int S, int C, sdvae_stream_t stream) {
    SDVAE_REQUIRE(W && Wd && (mode == 0 || mode == 1) && S > 0 && C > 0 && S * C <= 32 && N > 0 && N <= 32, "slot_weight: bad argument");
    slot_weight_kernel<<<4, 256, 0, (cudaStream_t)stream>>>(W, Wd, mode, N, S, C);
    return check_launch("slot_weight_kernel");
}

int sdvae_slot_grad(const float* dWd, const float* dbd, float* dW, float* db, int mode, int N, int S, int C,
                    sdvae_stream_t stream) {
    SDVAE_REQUIRE(dWd && dW && (mode == 0 || mode == 1) && S > 0 && C > 0 && S * C <= 32 && N > 0 && N <= 32, "slot_grad: bad argument");
    SDVAE_REQUIRE(!db || dbd, "slot_grad: db without dbd");
    slot_grad_kernel<<<4, 256, 0, (cudaStream_t)stream>>>(dWd, dbd, dW, db, mode, N, S, C);
    return check_launch("slot_grad_kernel");
}

int sdvae_tc_bwd_w_supported(int S, int Cin, int Cout, int rcap) {
    // C_in in 32-channel passes, C_out in passes of <= 32; 32*S + 1 accumulator rows in <= 3 blocks of 128
    if ((Cin != 32 && Cin != 64) || Cout < 1 || Cout > 64 || S < 1 || S > 11) return 0;
    if (rcap < 32 || rcap > umma::kMaxRcap || rcap % 32 != 0) return 0;
    const long long budget = 226LL * 1024 - 2048 - 2 * umma::kGStage;
    return budget / ((long long)rcap * 128) >= 7 ? 1 : 0;       // the phase-distance argument needs 7 raw stages
}

int sdvae_spiralconv_bwd_w_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src, int rcap,
                              const float* dpre, float* dW, float* db, void* workspace, int B, int Vin,
                              int Vout, int S, int Cin, int Cout, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && plan_cnt && plan_src && dpre && dW && workspace, "spiralconv_bwd_w_tc: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0, "spiralconv_bwd_w_tc: bad shape");
    SDVAE_REQUIRE((long long)B * Vin < 2147483647LL, "spiralconv_bwd_w_tc: B*Vin exceeds int32 rows");
    SDVAE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "spiralconv_bwd_w_tc: x must be 16-byte aligned");
    if (!sdvae_tc_bwd_w_supported(S, Cin, Cout, rcap))
        return set_error(SDVAE_ERR_UNSUPPORTED, "spiralconv_bwd_w_tc: unsupported layer shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = S * Cin;
    if (B == 0) {
        cudaMemsetAsync(dW, 0, sizeof(float) * Cout * K, st);
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * Cout, st);
        return check_launch("bwd_w_tc memset");
    }
    umma::BwUmmaArgs a{};
    a.plan_cnt = plan_cnt; a.plan_src = plan_src;
    a.B = B; a.in_rows = Vin; a.out_rows = Vout; a.L = sdvae_tc_plan_tiles(Vout); a.S = S; a.rcap = rcap;
    a.nraw = umma::kMaxRaw;
    const long long ntiles = (long long)B * a.L;
    const int grid = ntiles < kNumSMs ? (int)ntiles : kNumSMs;
    float* part = static_cast<float*>(workspace);               // [grid, Cout, K]
    float* part_b = part + (size_t)grid * Cout * K;             // [grid, Cout]
    static int flush_tiles = 0;                // tiles per accumulator drain (tuning knob, default 2)
    if (!flush_tiles) {
        flush_tiles = tuning_env("SDVAE_BWW_FLUSH", 2);
        if (flush_tiles < 1) flush_tiles = 1;
    }
    a.flush = flush_tiles;
    a.in_ld = Cin; a.g_ld = Cout; a.part_ld = K; a.part_cta = Cout * K; a.partb_cta = Cout;
    cudaMemsetAsync(part, 0, sizeof(float) * (size_t)grid * (Cout * K + Cout), st);
    static DeviceOnce attr_done;       // the attribute is per device
    if (attr_done.first()) cudaFuncSetAttribute(umma::bw_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const size_t smem = 1024 + 2 * (size_t)umma::kGStage + (size_t)a.nraw * rcap * 128 + 1024;
    // one pass per (32 input channels, <= 32 output channels); the bias gradient comes from the first
    // input-channel pass only (its row of ones)
    for (int c0 = 0; c0 < Cin; c0 += 32)
        for (int n0 = 0; n0 < Cout; n0 += umma::kBwNT) {
            a.in = x + c0;
            a.g = dpre + n0;
            a.n_real = Cout - n0 < umma::kBwNT ? Cout - n0 : umma::kBwNT;
            a.part = part + (size_t)n0 * K + c0;
            a.part_b = c0 == 0 ? part_b + n0 : nullptr;
            umma::bw_umma_kernel<<<grid, umma::kBwThreads, smem, st>>>(a);
            int rc = check_launch("bw_umma_kernel");
            if (rc) return rc;
        }
    const long long len = (long long)Cout * K;
    split_reduce2_kernel<<<blocks_for(len + Cout, 64), 64 * kSplitGroups, 0, st>>>(part, part_b, dW, db, grid, len, Cout);
    return check_launch("split_reduce_kernel");
}


/* Weight gradient with tile-local staging (spiral_conv_tile_bw.cuh): sdvae_spiralconv_bwd_w_tc on the FORWARD tile
 * plan of sdvae_spiralconv_fwd_tile (same workspace size). */
int sdvae_tile_bwd_w_supported(int S, int Cin, int Cout, int rcap) {
    if ((Cin != 32 && Cin != 64) || Cout < 1 || Cout > 64 || S < 1 || S * 32 + 1 > tile::kWMaxBlocks * 128) return 0;
    if (rcap < 32 || rcap > tile::kTMaxRcap || rcap % 32 != 0) return 0;
    return tile::TileBwCfg::stages(S, rcap) >= 2 ? 1 : 0;
}

int sdvae_spiralconv_bwd_w_tile(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                                const uint32_t* plan_cell, int rcap, const float* dpre, float* dW, float* db,
                                void* workspace, int B, int Vin, int Vout, int S, int Cin, int Cout,
                                sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && plan_cnt && plan_src && plan_cell && dpre && dW && workspace, "spiralconv_bwd_w_tile: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0, "spiralconv_bwd_w_tile: bad shape");
    SDVAE_REQUIRE((long long)B * Vin < 2147483647LL, "spiralconv_bwd_w_tile: B*Vin exceeds int32 rows");
    SDVAE_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(plan_src) |
                    reinterpret_cast<uintptr_t>(plan_cell)) & 15) == 0,
                  "spiralconv_bwd_w_tile: x and the plan tables must be 16-byte aligned");
    if (!sdvae_tile_bwd_w_supported(S, Cin, Cout, rcap))
        return set_error(SDVAE_ERR_UNSUPPORTED, "spiralconv_bwd_w_tile: unsupported layer shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = S * Cin;
    if (B == 0) {
        cudaMemsetAsync(dW, 0, sizeof(float) * Cout * K, st);
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * Cout, st);
        return check_launch("bwd_w_tile memset");
    }
    tile::TileBwArgs a{};
    a.plan_cnt = plan_cnt; a.plan_src = plan_src; a.plan_cell = plan_cell;
    a.B = B; a.in_rows = Vin; a.out_rows = Vout; a.L = sdvae_tc_plan_tiles(Vout); a.S = S; a.rcap = rcap;
    a.nts = tile::TileBwCfg::stages(S, rcap);
    const long long ntiles = (long long)B * a.L;
    const int grid = ntiles < kNumSMs ? (int)ntiles : kNumSMs;
    float* part = static_cast<float*>(workspace);               // [grid, Cout, K]
    float* part_b = part + (size_t)grid * Cout * K;             // [grid, Cout]
    static int flush_tiles = 0;                // tiles per accumulator drain (tuning knob, default 2)
    if (!flush_tiles) {
        flush_tiles = tuning_env("SDVAE_BWW_FLUSH", 2);
        if (flush_tiles < 1) flush_tiles = 1;
    }
    a.flush = flush_tiles;
    a.in_ld = Cin; a.g_ld = Cout; a.part_ld = K; a.part_cta = Cout * K; a.partb_cta = Cout;
    cudaMemsetAsync(part, 0, sizeof(float) * (size_t)grid * (Cout * K + Cout), st);
    cudaFuncSetAttribute(tile::bt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // per device, cheap
    const size_t smem = tile::TileBwCfg::smem_bytes(S, rcap, a.nts);
    for (int c0 = 0; c0 < Cin; c0 += 32)
        for (int n0 = 0; n0 < Cout; n0 += umma::kBwNT) {
            a.in = x + c0;
            a.g = dpre + n0;
            a.n_real = Cout - n0 < umma::kBwNT ? Cout - n0 : umma::kBwNT;
            a.part = part + (size_t)n0 * K + c0;
            a.part_b = c0 == 0 ? part_b + n0 : nullptr;
            tile::bt_kernel<<<grid, tile::kWThreads, smem, st>>>(a);
            int rc = check_launch("bt_kernel");
            if (rc) return rc;
        }
    const long long len = (long long)Cout * K;
    split_reduce2_kernel<<<blocks_for(len + Cout, 64), 64 * kSplitGroups, 0, st>>>(part, part_b, dW, db, grid, len, Cout);
    return check_launch("split_reduce_kernel");
}

size_t sdvae_spiralconv_bwd_w_workspace(long long M, int S, int Cin, int Cout) {
    long long rows; int nsplit;
    bw_split(M, &rows, &nsplit);
    if (nsplit < kNumSMs) nsplit = kNumSMs;      // the tensor-core path keeps one partial per CTA (<= one per SM)
    return (size_t)nsplit * (size_t)Cout * ((size_t)S * Cin + 1) * sizeof(float);
}

int sdvae_spiralconv_bwd_w(const float* x, const int32_t* idx, const float* dpre, float* dW,
                           float* db, void* workspace, int B, int Vin, int Vout, int S, int Cin,
                           int Cout, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && idx && dpre && dW && workspace, "spiralconv_bwd_w: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0 && S > 0 && Cin > 0 && Cout > 0, "spiralconv_bwd_w: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = S * Cin;
    BwArgs a{};
    a.in = x; a.idx = idx; a.g = dpre;
    a.M = (long long)B * Vout; a.in_rows = Vin; a.Vout = Vout; a.n_real = Cout;
    int nsplit;
    bw_split(a.M, &a.rows_per_cta, &nsplit);
    a.part = static_cast<float*>(workspace);
    a.part_b = a.part + (size_t)nsplit * Cout * K;
    int rc = SDVAE_OK;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dpre) & 15) == 0);
    if (a.M == 0) {
        cudaMemsetAsync(dW, 0, sizeof(float) * Cout * K, st);
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * Cout, st);
        return check_launch("bwd_w memset");
    }
    if (aligned && S == 9 && Cin == 32 && Cout == 32) rc = launch_bw<32, 9, 32, 8, 8>(a, nsplit, st);
    else if (aligned && S == 9 && Cin == 32 && Cout == 64) rc = launch_bw<32, 9, 64, 8, 8>(a, nsplit, st);
    else if (aligned && S == 9 && Cin == 64 && Cout == 64) rc = launch_bw<64, 9, 64, 8, 8>(a, nsplit, st);
    else if (aligned && S == 9 && Cin == 64 && Cout == 32) rc = launch_bw<64, 9, 32, 8, 8>(a, nsplit, st);
    else if (aligned && S == 9 && Cin == 32 && Cout == 3) rc = launch_bw<32, 9, 4, 2, 4>(a, nsplit, st);
    else if (S == 9 && Cin == 3 && Cout == 32 && aligned) rc = launch_bw<3, 9, 32, 2, 4>(a, nsplit, st);
    else {
        const dim3 grid((unsigned)nsplit, blocks_for((long long)Cout * (K + 1), 128));
        bw_generic_kernel<<<grid, 128, 0, st>>>(a, Cin, S);
        rc = check_launch("bw_generic_kernel");
    }
    if (rc) return rc;
    const long long len = (long long)Cout * K;
    split_reduce2_kernel<<<blocks_for(len + Cout, 64), 64 * kSplitGroups, 0, st>>>(a.part, a.part_b, dW, db, nsplit, len, Cout);
    return check_launch("split_reduce_kernel");
}

int sdvae_dense_fwd(const float* in, const float* W, const float* bias, float* out, long long M,
                    int K, int N, int ldw, int act, sdvae_stream_t stream) {
    SDVAE_REQUIRE(in && W && out && M >= 0 && K > 0 && N > 0 && ldw >= K, "dense_fwd: bad argument");
    SDVAE_REQUIRE(M < 2147483647LL, "dense_fwd: M exceeds int32 rows");
    GcArgs a{};
    a.in = in; a.idx = nullptr; a.W = W; a.bias = bias; a.out = out;
    a.M = M; a.in_rows = (int)(M > 0 ? M : 1); a.Vout = (int)(M > 0 ? M : 1); a.S = 1;
    a.ldw = ldw; a.ldo = N; a.n_real = N;
    const int epi = act == SDVAE_ACT_ELU ? EPI_BIAS_ELU : (bias ? EPI_BIAS : EPI_NONE);
    return dispatch_gc<false>(a, K, epi, (cudaStream_t)stream);
}

__global__ void transpose2d_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = in[(size_t)r * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

int sdvae_transpose2d(const float* in, float* out, int R, int C, sdvae_stream_t stream) {
    SDVAE_REQUIRE(in && out && R > 0 && C > 0, "transpose2d: bad argument");
    const dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    transpose2d_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, R, C);
    return check_launch("transpose2d_kernel");
}

int sdvae_pool_ell_fwd(const float* x, const int32_t* col, const float* val, float* out, int B,
                       int Vin, int Vout, int Wd, int C, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && col && val && out, "pool_ell_fwd: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0 && Wd > 0 && C > 0, "pool_ell_fwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const long long total = (long long)B * Vout * (vec ? C / 4 : C);
    if (total == 0) return SDVAE_OK;
    if (vec && B >= kPoolMeshes) {
        const long long tb = (long long)((B + kPoolMeshes - 1) / kPoolMeshes) * Vout * (C / 4);
        pool_ell_fwd_batch_kernel<<<blocks_for(tb, 256), 256, 0, st>>>(x, col, val, out, tb, B, Vin, Vout, Wd, C);
        return check_launch("pool_ell_fwd_batch_kernel");
    }
    if (vec) pool_ell_fwd_kernel<true><<<blocks_for(total, 256), 256, 0, st>>>(x, col, val, out, total, Vin, Vout, Wd, C);
    else pool_ell_fwd_kernel<false><<<blocks_for(total, 256), 256, 0, st>>>(x, col, val, out, total, Vin, Vout, Wd, C);
    return check_launch("pool_ell_fwd_kernel");
}

// Row-range parts per mesh for the kernels that keep a whole mesh resident in shared memory: every work item
// (mesh, part) stages the mesh again (stage_us) and processes rows/parts rows (row_ns each); the CTAs (one per
// SM) take ceil(B*parts/SMs) items each.  Measured on B200: staging 204 KB ~6 us; 13 ns per row for the 3 -> 32
// kernels, 22 ns per row for the fused 32 -> 3 backward (model error < 10 % over B = 64..1024, parts = 1..8).
static int narrow_parts(int B, int rows, double stage_us, double row_ns) {
    { static int env = -1; if (env < 0) env = tuning_env("SDVAE_NARROW_PARTS", 0);
      if (env > 0) return env; }
    int best = 1;
    double best_cost = 0.0;
    for (int p = 1; p <= 16 && rows / p >= 64; ++p) {
        const long long rounds = ((long long)B * p + kNumSMs - 1) / kNumSMs;
        const double cost = rounds * (stage_us + 1e-3 * row_ns * ((rows + p - 1) / p));
        if (p == 1 || cost < best_cost) { best_cost = cost; best = p; }
    }
    return best;
}
/* ---- narrow-output layer backward (narrow_conv.cuh) ------------------------------------------ */
int sdvae_narrow_out_bwd_supported(int R, int S, int Cin, int Cout) {
    if (!(S == 9 && Cout == 3 && Cin == 32 && R > 0 && R * Cout < 0xffff)) return 0;      // cell_pack: 16-bit offsets
    return NarrowCfg<9, 3>::smem_bytes(R) <= 227 * 1024 ? 1 : 0;
}

size_t sdvae_narrow_out_bwd_workspace(int S, int Cout) {
    return sizeof(float) * (size_t)kNumSMs * ((size_t)S * Cout * 32 + 32);
}

int sdvae_narrow_out_bwd(const float* dy, const float* x, const int32_t* cell_ptr, const int32_t* cell_src,
                         const int32_t* cell_pack, const float* W, float* dx, float* dW, float* db, void* workspace, int B, int R,
                         int Vin, int S, int Cin, int Cout, int gated, sdvae_stream_t stream) {
    SDVAE_REQUIRE(dy && x && cell_ptr && cell_src && cell_pack && W && workspace, "narrow_out_bwd: null pointer");
    SDVAE_REQUIRE((reinterpret_cast<uintptr_t>(cell_pack) & 7) == 0, "narrow_out_bwd: cell_pack must be 8-byte aligned");
    SDVAE_REQUIRE(B >= 0 && Vin > 0, "narrow_out_bwd: bad shape");
    SDVAE_REQUIRE(sdvae_narrow_out_bwd_supported(R, S, Cin, Cout), "narrow_out_bwd: unsupported shape (see sdvae_narrow_out_bwd_supported)");
    cudaStream_t st = (cudaStream_t)stream;
    using Cfg = NarrowCfg<9, 3>;
    float* part = static_cast<float*>(workspace);
    int grid = 0;
    if (B > 0) {
        auto kern = narrow_out_bwd_kernel<9, 3>;
        static DeviceOnce attr_done;       // the attribute is per device
        if (attr_done.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        const int parts = narrow_parts(B, Vin, 6.0 * R / 17039.0, 22.0);
        const long long items = (long long)B * parts;
        grid = items < kNumSMs ? (int)items : kNumSMs;
        kern<<<grid, kNarrowThreads, Cfg::smem_bytes(R), st>>>(dy, x, cell_ptr, cell_src, cell_pack, W, dx, part, B, parts, R,
                                                               Vin, gated, (int)Cfg::main_floats(R));
        const int rc = check_launch("narrow_out_bwd_kernel");
        if (rc) return rc;
    }
    if (dW || db) {
        narrow_out_reduce_kernel<<<blocks_for(32LL * Cfg::PART, 256), 256, 0, st>>>(part, grid, dW, db, S, Cout);
        return check_launch("narrow_out_reduce_kernel");
    }
    return SDVAE_OK;
}

/* ---- narrow-input layer (3 -> 32): forward and weight gradient (narrow_conv.cuh) ------------------ */
int sdvae_narrow_in_supported(int Vin, int S, int Cin, int Cout) {
    if (!(S == 9 && Cin == 3 && Cout == 32 && Vin > 0)) return 0;
    return NarrowInCfg<9, 3>::smem_bytes(Vin) <= 227 * 1024 ? 1 : 0;
}

size_t sdvae_narrow_in_bwd_w_workspace(int S, int Cin) {
    return sizeof(float) * (size_t)kNumSMs * ((size_t)S * Cin * 32 + 32);
}

static int narrow_in_parts(int B, int R, int Vin) { return narrow_parts(B, R, 6.0 * Vin / 17039.0, 13.0); }

int sdvae_narrow_in_fwd(const float* x, const int32_t* idx, const float* W, const float* bias, float* y, int B,
                        int Vin, int R, int S, int Cin, int Cout, int act, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && idx && W && y, "narrow_in_fwd: null pointer");
    SDVAE_REQUIRE(B >= 0 && R > 0 && (act == SDVAE_ACT_NONE || act == SDVAE_ACT_ELU), "narrow_in_fwd: bad argument");
    SDVAE_REQUIRE(sdvae_narrow_in_supported(Vin, S, Cin, Cout), "narrow_in_fwd: unsupported shape (see sdvae_narrow_in_supported)");
    if (B == 0) return SDVAE_OK;
    using Cfg = NarrowInCfg<9, 3>;
    auto kern = narrow_in_kernel<9, 3, 0>;
    static DeviceOnce attr_done;       // the attribute is per device
    if (attr_done.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int parts = narrow_in_parts(B, R, Vin);
    const long long items = (long long)B * parts;
    const int grid = items < kNumSMs ? (int)items : kNumSMs;
    kern<<<grid, kNarrowThreads, Cfg::smem_bytes(Vin), (cudaStream_t)stream>>>(x, idx, W, bias, y, B, parts, R, Vin, act,
                                                                               (int)Cfg::main_floats(Vin));
    return check_launch("narrow_in_kernel<fwd>");
}

int sdvae_narrow_in_bwd_w(const float* x, const int32_t* idx, const float* dpre, float* dW, float* db,
                          void* workspace, int B, int Vin, int R, int S, int Cin, int Cout,
                          sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && idx && dpre && workspace, "narrow_in_bwd_w: null pointer");
    SDVAE_REQUIRE(B >= 0 && R > 0, "narrow_in_bwd_w: bad shape");
    SDVAE_REQUIRE(sdvae_narrow_in_supported(Vin, S, Cin, Cout), "narrow_in_bwd_w: unsupported shape (see sdvae_narrow_in_supported)");
    using Cfg = NarrowInCfg<9, 3>;
    cudaStream_t st = (cudaStream_t)stream;
    float* part = static_cast<float*>(workspace);
    int grid = 0;
    if (B > 0) {
        auto kern = narrow_in_kernel<9, 3, 1>;
        static DeviceOnce attr_done;       // the attribute is per device
        if (attr_done.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        const int parts = narrow_in_parts(B, R, Vin);
        const long long items = (long long)B * parts;
        grid = items < kNumSMs ? (int)items : kNumSMs;
        kern<<<grid, kNarrowThreads, Cfg::smem_bytes(Vin), st>>>(x, idx, nullptr, dpre, part, B, parts, R, Vin, 0,
                                                                 (int)Cfg::main_floats(Vin));
        const int rc = check_launch("narrow_in_kernel<bwd_w>");
        if (rc) return rc;
    }
    narrow_in_reduce_kernel<<<blocks_for(32LL * Cfg::PART, 256), 256, 0, st>>>(part, grid, dW, db, Cfg::J);
    return check_launch("narrow_in_reduce_kernel");
}

int sdvae_narrow_out_fwd_tile(void) { return kNarrowTile; }

int sdvae_narrow_out_fwd_supported(int S, int Cin, int Cout, int ucap) {
    return S == 9 && Cin == 32 && Cout == 3 && ucap > 0 && ucap * 8 <= kNarrowMaxIssue * kNarrowFwdThreads &&
           (long long)kNarrowMinStages * ucap * 128 <= 220 * 1024 ? 1 : 0;
}

int sdvae_narrow_out_fwd(const float* x, const int32_t* tile_ptr, const int32_t* stage_src, const int32_t* loc,
                         const float* W, const float* bias, float* out, int B, int Vin, int Vout, int S, int Cin,
                         int Cout, int T, int ucap, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && tile_ptr && stage_src && loc && W && out, "narrow_out_fwd: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0, "narrow_out_fwd: bad shape");
    SDVAE_REQUIRE(T == kNarrowTile, "narrow_out_fwd: the plan's tile must be sdvae_narrow_out_fwd_tile() rows");
    SDVAE_REQUIRE(sdvae_narrow_out_fwd_supported(S, Cin, Cout, ucap), "narrow_out_fwd: unsupported shape (see sdvae_narrow_out_fwd_supported)");
    SDVAE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "narrow_out_fwd: x must be 16-byte aligned");
    if (B == 0) return SDVAE_OK;
    int nst = (int)((220LL * 1024) / ((long long)ucap * 128));
    nst = std::max(kNarrowMinStages, std::min(kNarrowMaxStages, nst));
    { static int env = -1; if (env < 0) env = tuning_env("SDVAE_NARROW_STAGES", 0);
      if (env >= kNarrowMinStages && env <= nst) nst = env; }
    auto kern = nst == 5 ? narrow_out_fwd_kernel<9, 3, 5> : nst == 4 ? narrow_out_fwd_kernel<9, 3, 4> : narrow_out_fwd_kernel<9, 3, 3>;
    static DeviceOnce attr_done[8];      // per stage count and per device
    if (attr_done[nst].first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const size_t smem = (size_t)nst * ucap * 128;
    const int L = (Vout + kNarrowTile - 1) / kNarrowTile;
    const int MG = pick_mesh_group(B, L, kNumSMs, nst);          // one CTA per SM (registers)
    const long long grid = (long long)L * ((B + MG - 1) / MG);
    kern<<<(unsigned)grid, kNarrowFwdThreads, smem, (cudaStream_t)stream>>>(x, tile_ptr, stage_src, loc, W, bias, out,
                                                                            B, Vin, Vout, L, ucap, MG);
    return check_launch("narrow_out_fwd_kernel");
}

int sdvae_pool_stage_tile(void) { return kPoolTile; }

int sdvae_pool_stage_supported(int C, int Wd, int ucap) {
    return (C == 32 || C == 64) && Wd >= 2 && Wd <= 4 && ucap > 0 &&
           ucap * (C / 4) <= kPoolMaxIssue * kPoolStageThreads &&
           (long long)pool_stages_for(C / 4) * ucap * C * 4 <= 200 * 1024 ? 1 : 0;
}

int sdvae_pool_ell_fwd_staged(const float* x, const int32_t* tile_ptr, const int32_t* stage_src,
                              const int32_t* ent, float* out, int B, int Vin, int Vout, int Wd, int C,
                              int T, int ucap, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && tile_ptr && stage_src && ent && out, "pool_ell_fwd_staged: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vin > 0 && Vout > 0, "pool_ell_fwd_staged: bad shape");
    SDVAE_REQUIRE(T == kPoolTile, "pool_ell_fwd_staged: the plan's tile must be sdvae_pool_stage_tile() rows");
    SDVAE_REQUIRE(sdvae_pool_stage_supported(C, Wd, ucap), "pool_ell_fwd_staged: unsupported C / width / staged rows per tile (see sdvae_pool_stage_supported)");
    SDVAE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(ent) & 7) == 0, "pool_ell_fwd_staged: misaligned pointer");
    if (B == 0) return SDVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 32) return launch_pool_staged_w<8>(x, tile_ptr, stage_src, ent, out, B, Vin, Vout, Wd, ucap, st);
    return launch_pool_staged_w<16>(x, tile_ptr, stage_src, ent, out, B, Vin, Vout, Wd, ucap, st);
}

int sdvae_csr_rowsum(const float* dy, const int32_t* ptr, const int32_t* src, const float* val,
                     const float* gate, float* dx, int B, int Vsrc, int Vdst, int C,
                     sdvae_stream_t stream) {
    SDVAE_REQUIRE(dy && ptr && src && dx, "csr_rowsum: null pointer");
    SDVAE_REQUIRE(B >= 0 && Vsrc > 0 && Vdst > 0 && C > 0, "csr_rowsum: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dx) & 15) == 0) &&
                     (!gate || (reinterpret_cast<uintptr_t>(gate) & 15) == 0);
    const long long total = (long long)B * Vdst * (vec ? C / 4 : C);
    if (total == 0) return SDVAE_OK;
    if (vec && (C == 32 || C == 64) && B >= 2 * (128 / C)) {
        // one output row per warp (uniform trip count): C/4 lanes x 128/C meshes, kPoolMeshes more in registers
        const int MW = (128 / C) * kPoolMeshes;
        const long long warps = (long long)((B + MW - 1) / MW) * Vdst;
        if (C == 32) csr_rowsum_warp_kernel<8><<<blocks_for(warps * 32, 256), 256, 0, st>>>(dy, ptr, src, val, gate, dx, warps, B, Vsrc, Vdst);
        else csr_rowsum_warp_kernel<16><<<blocks_for(warps * 32, 256), 256, 0, st>>>(dy, ptr, src, val, gate, dx, warps, B, Vsrc, Vdst);
        return check_launch("csr_rowsum_warp_kernel");
    }
    if (vec && B >= kPoolMeshes) {
        const long long tb = (long long)((B + kPoolMeshes - 1) / kPoolMeshes) * Vdst * (C / 4);
        csr_rowsum_batch_kernel<<<blocks_for(tb, 256), 256, 0, st>>>(dy, ptr, src, val, gate, dx, tb, B, Vsrc, Vdst, C);
        return check_launch("csr_rowsum_batch_kernel");
    }
    if (vec) csr_rowsum_kernel<true><<<blocks_for(total, 256), 256, 0, st>>>(dy, ptr, src, val, gate, dx, total, Vsrc, Vdst, C);
    else csr_rowsum_kernel<false><<<blocks_for(total, 256), 256, 0, st>>>(dy, ptr, src, val, gate, dx, total, Vsrc, Vdst, C);
    return check_launch("csr_rowsum_kernel");
}

int sdvae_elu_fwd(const float* x, float* y, long long n, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && y && n >= 0, "elu_fwd: bad argument");
    if (n == 0) return SDVAE_OK;
    elu_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    return check_launch("elu_kernel");
}

int sdvae_elu_bwd(const float* dy, const float* y, float* dx, long long n, sdvae_stream_t stream) {
    SDVAE_REQUIRE(dy && y && dx && n >= 0, "elu_bwd: bad argument");
    if (n == 0) return SDVAE_OK;
    elu_gate_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n);
    return check_launch("elu_gate_kernel");
}

int sdvae_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z,
                      long long n, sdvae_stream_t stream) {
    SDVAE_REQUIRE(mu && logvar && eps && z && n >= 0, "reparam_fwd: bad argument");
    if (n == 0) return SDVAE_OK;
    reparam_fwd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(mu, logvar, eps, z, n);
    return check_launch("reparam_fwd_kernel");
}

int sdvae_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu,
                      float* dlogvar, long long n, int accumulate, sdvae_stream_t stream) {
    SDVAE_REQUIRE(dz && logvar && eps && dmu && dlogvar && n >= 0, "reparam_bwd: bad argument");
    if (n == 0) return SDVAE_OK;
    reparam_bwd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dz, logvar, eps, dmu, dlogvar, n, accumulate);
    return check_launch("reparam_bwd_kernel");
}

int sdvae_axpy3(const float* a, const float* b, float sb, const float* c, float sc, float* out,
                long long n, sdvae_stream_t stream) {
    SDVAE_REQUIRE(out && n >= 0, "axpy3: bad argument");
    if (n == 0) return SDVAE_OK;
    axpy3_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, sb, c, sc, out, n);
    return check_launch("axpy3_kernel");
}

int sdvae_swap(const float* x, const uint8_t* mask, float* out, int bs, int i0, int i1, int V,
               int C, sdvae_stream_t stream) {
    SDVAE_REQUIRE(x && mask && out, "swap: null pointer");
    SDVAE_REQUIRE(bs > 0 && i0 >= 0 && i1 >= i0 && i1 <= bs && V > 0 && C > 0, "swap: bad shape");
    const long long total = (long long)(i1 - i0) * bs * V * C;
    if (total == 0) return SDVAE_OK;
    swap_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, mask, out, bs, i0, i1, V, C);
    return check_launch("swap_kernel");
}

size_t sdvae_mse_lap_partial_floats(int B, int V) {
    return 2 * (size_t)blocks_for((long long)B * V, kLossThreads);
}

int sdvae_mse_lap_fwd(const float* recon, const float* x, const int32_t* lcol, const float* lval,
                      int lw, float* qn, float* partial, float* losses, int B, int V,
                      float inv_count_scale, sdvae_stream_t stream) {
    SDVAE_REQUIRE(recon && x && partial && losses && B > 0 && V > 0, "mse_lap_fwd: bad argument");
    SDVAE_REQUIRE(!lcol || (lval && qn && lw > 0), "mse_lap_fwd: Laplacian tables incomplete");
    cudaStream_t st = (cudaStream_t)stream;
    const long long BV = (long long)B * V;
    const unsigned grid = blocks_for(BV, kLossThreads);
    mse_lap_fwd_kernel<<<grid, kLossThreads, 0, st>>>(recon, x, lcol, lval, lw, qn, partial, BV, V);
    finish_sum_kernel<<<1, kLossThreads, 0, st>>>(partial, grid, 2, 0, inv_count_scale / (float)((double)BV * 3.0), losses, 0);
    if (lcol) finish_sum_kernel<<<1, kLossThreads, 0, st>>>(partial, grid, 2, 1, inv_count_scale / (float)BV, losses, 3);
    return check_launch("mse_lap_fwd");
}

int sdvae_l1_fwd(const float* a, const float* b, float* partial, float* out, long long n, sdvae_stream_t stream) {
    SDVAE_REQUIRE(a && b && partial && out && n > 0, "l1_fwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = blocks_for(n, kLossThreads);
    l1_fwd_kernel<<<grid, kLossThreads, 0, st>>>(a, b, partial, n);
    finish_sum_kernel<<<1, kLossThreads, 0, st>>>(partial, grid, 1, 0, 1.0f / (float)n, out, 0);
    return check_launch("l1_fwd");
}

int sdvae_l1_bwd(const float* a, const float* b, float* da, long long n, float g, sdvae_stream_t stream) {
    SDVAE_REQUIRE(a && b && da && n > 0, "l1_bwd: bad argument");
    l1_bwd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, da, n, g / (float)n);
    return check_launch("l1_bwd_kernel");
}

int sdvae_mse_lap_bwd(const float* recon, const float* x, const float* qn, const int32_t* tptr,
                      const int32_t* trow, const float* tval, float* drecon, int B, int V,
                      float g_mse, float g_lap, float inv_count_scale, const float* dscale,
                      sdvae_stream_t stream) {
    SDVAE_REQUIRE(recon && x && drecon && B > 0 && V > 0, "mse_lap_bwd: bad argument");
    SDVAE_REQUIRE(!tptr || (trow && tval && qn), "mse_lap_bwd: Laplacian tables incomplete");
    const long long BV = (long long)B * V;
    const float c_mse = g_mse * inv_count_scale / (float)((double)BV * 3.0);
    const float c_lap = g_lap * inv_count_scale / (float)BV;
    mse_lap_bwd_kernel<<<blocks_for(BV, 256), 256, 0, (cudaStream_t)stream>>>(
        recon, x, qn, tptr, trow, tval, drecon, BV, V, c_mse, c_lap, dscale);
    return check_launch("mse_lap_bwd_kernel");
}

int sdvae_kl_fwd_bwd(const float* mu, const float* logvar, float* dmu, float* dlogvar,
                     float* partial, float* losses, int B, int D, float inv_count_scale,
                     sdvae_stream_t stream) {
    SDVAE_REQUIRE(mu && logvar && dmu && dlogvar && partial && losses && B > 0 && D > 0, "kl: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * D;
    const unsigned grid = blocks_for(n, kLossThreads);
    const float inv_b = inv_count_scale / (float)B;
    kl_fwd_bwd_kernel<<<grid, kLossThreads, 0, st>>>(mu, logvar, dmu, dlogvar, partial, n, inv_b);
    finish_sum_kernel<<<1, kLossThreads, 0, st>>>(partial, grid, 1, 0, inv_b, losses, 1);
    return check_launch("kl_fwd_bwd");
}

int sdvae_lc_fwd_bwd(const float* z, int bs, int D, int r0, int r1, float eta1, float eta2,
                     uint8_t* act_ws, float* partial, float* dz, float* losses,
                     sdvae_stream_t stream) {
    SDVAE_REQUIRE(z && act_ws && partial && dz && losses, "lc: null pointer");
    SDVAE_REQUIRE(bs >= 2 && D > 0 && r0 >= 0 && r1 > r0 && r1 <= D, "lc: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nh = (long long)bs * (bs - 1) / 2 * bs;
    const unsigned grid = blocks_for(nh, kLossThreads);
    const double norm = (double)bs * bs * bs - (double)bs * bs;
    lc_hinge_kernel<<<grid, kLossThreads, 0, st>>>(z, bs, D, r0, r1, eta1, eta2, act_ws, partial);
    finish_sum_kernel<<<1, kLossThreads, 0, st>>>(partial, grid, 1, 0, (float)(1.0 / norm), losses, 2);
    const long long ng = (long long)bs * bs * D;
    lc_grad_kernel<<<blocks_for(ng, 256), 256, 0, st>>>(z, act_ws, bs, D, r0, r1, (float)(2.0 / norm), dz);
    return check_launch("lc_fwd_bwd");
}

int sdvae_total_loss(float* losses, float w_kl, float w_lc, float w_lap, float w_cls,
                     sdvae_stream_t stream) {
    SDVAE_REQUIRE(losses, "total_loss: null pointer");
    total_loss_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(losses, w_kl, w_lc, w_lap, w_cls);
    return check_launch("total_loss_kernel");
}

int sdvae_adam_tick(int32_t* step_dev, sdvae_stream_t stream) {
    SDVAE_REQUIRE(step_dev, "adam_tick: null pointer");
    adam_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_dev);
    return check_launch("adam_tick_kernel");
}

int sdvae_adam_step(float* p, const float* grad, float* m, float* v, long long n,
                    const int32_t* step_dev, int step_host, float lr, float beta1, float beta2,
                    float eps, float weight_decay, float gscale, sdvae_stream_t stream) {
    SDVAE_REQUIRE(p && grad && m && v && n >= 0, "adam_step: bad argument");
    SDVAE_REQUIRE(step_dev || step_host >= 1, "adam_step: step must be >= 1");
    if (n == 0) return SDVAE_OK;
    adam_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        p, grad, m, v, n, step_dev, step_host, lr, beta1, beta2, eps, weight_decay, gscale);
    return check_launch("adam_kernel");
}

}  // extern "C"
