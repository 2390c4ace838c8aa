/* sdvae_b200 -- C ABI of the B200-native SD-VAE mesh encoder/decoder hot path.
 *
 * Drop-in boundary for the reference's `model.py` operators (SpiralConv, Pool) and the
 * loss / optimiser arithmetic of `model_manager.py::_do_iteration`.  The reference is pure
 * PyTorch, so "what its FFI would bind" is one entry point per ATen op group it calls on the
 * path; each declaration cites the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (fp32 / int32 / uint8); tensors are dense row-major
 *   - the caller allocates every output and workspace; nothing is allocated or cached inside
 *   - no thread-local or global state except the last-error string; safe to call from the
 *     autograd engine thread; the CUDA device must be current, `stream` is a cudaStream_t
 *   - return 0 on success, non-zero on error (1 bad argument, 2 CUDA error, 3 unsupported);
 *     sdvae_last_error() then describes it.  There is no CPU fallback.
 *   - index tables are int32 (the reference stores int64 with values < 2^31; range-checked on
 *     the host when the tables are derived, never at kernel time)
 */
#ifndef SDVAE_B200_H
#define SDVAE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sdvae_stream_t;   /* cudaStream_t */

#define SDVAE_ACT_NONE 0
#define SDVAE_ACT_ELU  1

int         sdvae_abi_version(void);
const char* sdvae_last_error(void);

/* ---- SpiralConv ------------------------------------------------------------------------- */

/* y[b,v,o] = act(bias[o] + sum_{s,c} W[o, s*Cin+c] * x[b, idx[v,s], c])
 * x [B,Vin,Cin], idx [Vout,S] (values in [0,Vin)), W [Cout, S*Cin] (nn.Linear layout),
 * bias [Cout] or NULL, y [B,Vout,Cout].  Vout may be a SUBSET of the vertices (rows of the
 * spiral table restricted to the vertices a selection down-transform keeps), which fuses
 * `Pool(elu(conv(x)), down)`.
 * Replaces: model.py:27-41 (index_select + nn.Linear) and F.elu at model.py:68,84. */
int sdvae_spiralconv_fwd(const float* x, const int32_t* idx, const float* W, const float* bias,
                         float* y, int B, int Vin, int Vout, int S, int Cin, int Cout, int act,
                         sdvae_stream_t stream);

/* Wt[c, s*Cout+o] = W[o, s*Cin+c]: weight of the transposed convolution used by bwd_x. */
int sdvae_weight_transpose(const float* W, float* Wt, int Cout, int Cin, int S,
                           sdvae_stream_t stream);

/* dx[b,u,c] = gate * sum_{s} sum_{r in cell(u,s)} sum_o dpre[b,r,o] * W[o, s*Cin+c]
 * dpre [B,Vrows,Cout] (gradient w.r.t. the pre-activation), cell_ptr [Vdst*S+1] / cell_src:
 * for input vertex u and slot s the ascending list of output rows r with idx[r,s]==u,
 * Wt from sdvae_weight_transpose, gate [B,Vdst,Cin] or NULL (if given, dx *= elu'(gate), i.e.
 * the ELU derivative of the layer that produced x, evaluated from its output), dx [B,Vdst,Cin].
 * Deterministic: fixed summation order, no atomics.
 * Replaces: autograd of model.py:34 (index_select backward = index_add_ with atomics) and of
 * model.py:40 (grad_input GEMM). */
int sdvae_spiralconv_bwd_x(const float* dpre, const int32_t* cell_ptr, const int32_t* cell_src,
                           const float* Wt, const float* gate, float* dx, int B, int Vrows,
                           int Vdst, int S, int Cout, int Cin, sdvae_stream_t stream);

/* ---- SpiralConv on the 5th-gen tensor cores (tcgen05.mma + TMEM) ------------------------------
 * Same contractions as sdvae_spiralconv_fwd / sdvae_spiralconv_bwd_x, for the wide layers
 * (channels per slot KS in {32, 64}; outputs N <= 64).
 * Arithmetic: error-compensated 3xTF32 (hi/lo operand split, fp32 accumulation in TMEM) -- the
 * parity bar of this path is stated separately in tests/ (normwise vs the fp64 oracle).
 *
 * Weight operand: a packed image (already split and laid out as the 128B-swizzled UMMA tiles)
 * produced by sdvae_tc_pack_weights from the nn.Linear weight W [Cout, S*Cin] (model.py:16-21);
 * `transposed` != 0 packs the backward-to-input operand Wt[c, s*Cout+o] = W[o, s*Cin+c].
 * Re-pack whenever W changes.
 *
 * Gather operand: a TILE PLAN built once per index table on the HOST (sdvae_tc_plan_* take host
 * pointers) and then copied to the device.  The table is given in cell form: for output row r and
 * slot s the source rows cell_src[cell_ptr[r*S+s] .. cell_ptr[r*S+s+1]) whose SUM feeds the
 * contraction.  The forward table idx [out_rows, S] (model.py:18, spirals.pkl) is the cell form
 * with cell_ptr[i] = i, cell_src = idx; the backward table is the inverse (cell_ptr, cell_src)
 * of sdvae_spiralconv_bwd_x.  Plan arrays, L = sdvae_tc_plan_tiles(out_rows) tiles of 128 rows:
 *   cnt  [L, S]        rows staged for (tile, slot)
 *   src  [L, S, rcap/2] their source rows (< 65536), two per 32-bit word, packed in the order the
 *                      loader lanes consume them; rcap = sdvae_tc_plan_max_rows(...) rounded up to 32
 *   cell [L, S, 128]   start | count << 16 : staged rows summed (in order) into each tile row */
int    sdvae_tc_supported(int S, int KS, int N, int rcap);
size_t sdvae_tc_wimg_floats(int S, int KS, int N);
int sdvae_tc_pack_weights(const float* W, float* wimg, int S, int Cin, int Cout, int transposed,
                          sdvae_stream_t stream);
/* Several weight images in ONE launch (the images of a training step are re-packed every step; one launch
 * instead of ~20 three-microsecond ones).  `entries` is a DEVICE array of n sdvae_pack_entry. */
typedef struct {
    const float* W;     /* layer weight [Cout, S*Cin] */
    float* wimg;        /* destination image, sdvae_tc_wimg_floats(S, KS, n_cnt) floats */
    int S, Cin, Cout, transposed /* flag word: bit 0 transposed, bit 1 kperm image */, n0, n_cnt;
} sdvae_pack_entry;
int sdvae_tc_pack_weights_batch(const sdvae_pack_entry* entries, int n, sdvae_stream_t stream);
/* The same for output channels [n0, n0 + n_cnt) only (input channels for the transposed weight): a layer
 * whose full weight image does not fit in shared memory (64 -> 64) runs as two 32-channel passes, each
 * writing its columns of the output through ldy / lddx below. */
int sdvae_tc_pack_weights_part(const float* W, float* wimg, int S, int Cin, int Cout, int transposed,
                               int n0, int n_cnt, sdvae_stream_t stream);
int sdvae_tc_plan_tiles(int out_rows);
int sdvae_tc_plan_max_rows(const int32_t* cell_ptr, int out_rows, int S);
int sdvae_tc_plan_build(const int32_t* cell_ptr, const int32_t* cell_src, int out_rows, int S,
                        int rcap, int32_t* cnt, int32_t* src, int32_t* cell);
/* Replaces: model.py:27-41 + F.elu (model.py:68,84), as sdvae_spiralconv_fwd.  The plan must be the
 * FORWARD plan of the layer's table (cell_ptr[i] = i: exactly one source row per cell, so staged row e of a
 * (tile, slot) is tile row e); plan_cell is not read and may be NULL.  ldy: floats between output rows
 * (0 = Cout); with ldy > Cout, y points at the first of Cout consecutive columns of a wider tensor. */
int sdvae_spiralconv_fwd_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                            const int32_t* plan_cell, int rcap, const float* wimg, const float* bias,
                            float* y, int B, int Vin, int Vout, int S, int Cin, int Cout, int act, int ldy,
                            sdvae_stream_t stream);
/* Replaces: autograd of model.py:34,40 w.r.t. the input, as sdvae_spiralconv_bwd_x. */
int sdvae_spiralconv_bwd_x_tc(const float* dpre, const int32_t* plan_cnt, const int32_t* plan_src,
                              const int32_t* plan_cell, int rcap, const float* wimg_t, const float* gate,
                              float* dx, int B, int Vrows, int Vdst, int S, int Cout, int Cin, int lddx,
                              sdvae_stream_t stream);

/* The same two passes with TILE-LOCAL STAGING (csrc/spiral_conv_tile.cuh), for 32 -> 32 channel layers whose
 * level is numbered so that a tile of 128 consecutive output rows reads at most 288 distinct source rows
 * (tables.patch_order): the tile's distinct rows are copied once into shared memory and gathered there, the
 * A operand reaches the MMA thread in stages of three K chunks.  Replaces model.py:27-41 + F.elu
 * (model.py:68,84) / autograd of model.py:34,40 w.r.t. the input, exactly as the two entries above.
 * Tile plan (tables.tile_plan), per tile of 128 output rows:
 *   plan_cnt  [L]           distinct source rows of the tile
 *   plan_src  [L, rcap/2]   those rows, 16-bit pairs in loader-lane order (as sdvae_tc_plan_build's src with S = 1)
 *   plan_cell [L, S*128]    word of (slot s, tile row r) at s*128 + (r>>5)*32 + (r&7)*4 + ((r>>3)&3):
 *                           bits 0..15 byte offset of the low 64-byte half of the cell's first row in the tile stage
 *                           (position p: p*128 + 64*(p&1): rows at odd positions are stored high half first),
 *                           bits 16..20 rows in the cell, bits 21.. offset of the cell's further rows in plan_ext
 *   plan_ext  [L, ecap]     (backward plans) 16-bit byte offsets (same form) of the 2nd, 3rd ... rows of the cells
 * wimg: sdvae_tc_pack_weights with flag bit 1 set (`transposed | 2`: K positions of every 32-wide chunk permuted
 * the way the kernel's conflict-free shared-memory gather delivers them). */
int sdvae_tile_supported(int S, int Cin, int Cout, int rcap, int ecap);
int sdvae_spiralconv_fwd_tile(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                              const uint32_t* plan_cell, int rcap, const float* wimg, const float* bias, float* y,
                              int B, int Vin, int Vout, int S, int Cin, int Cout, int act, sdvae_stream_t stream);
int sdvae_spiralconv_bwd_x_tile(const float* dpre, const int32_t* plan_cnt, const int32_t* plan_src,
                                const uint32_t* plan_cell, const uint16_t* plan_ext, int rcap, int ecap,
                                const float* wimg_t, const float* gate, float* dx, int B, int Vrows, int Vdst, int S,
                                int Cout, int Cin, sdvae_stream_t stream);

/* Weight gradient with tile-local staging (csrc/spiral_conv_tile_bw.cuh): as sdvae_spiralconv_bwd_w_tc below (same
 * contraction, same workspace size, same deterministic drain), on the FORWARD tile plan of sdvae_spiralconv_fwd_tile.
 * Replaces: autograd of nn.Linear in model.py:40 over the gather of model.py:34.  C_in in {32, 64}, C_out <= 64. */
int sdvae_tile_bwd_w_supported(int S, int Cin, int Cout, int rcap);
int sdvae_spiralconv_bwd_w_tile(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                                const uint32_t* plan_cell, int rcap, const float* dpre, float* dW, float* db,
                                void* workspace, int B, int Vin, int Vout, int S, int Cin, int Cout,
                                sdvae_stream_t stream);

/* Weight gradient on the tensor cores, as sdvae_spiralconv_bwd_w (same workspace size), for
 * C_in in {32, 64}, C_out <= 64 (passes of 32 x <= 32 channels).  (plan_cnt, plan_src, rcap) is the FORWARD tile plan of the layer's table
 * (one source row per cell).  db comes from a row of ones in the A operand; partial sums per CTA are
 * added in a fixed order.  Replaces: autograd of model.py:40 (grad_weight / grad_bias). */
int sdvae_tc_bwd_w_supported(int S, int Cin, int Cout, int rcap);
int sdvae_spiralconv_bwd_w_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src, int rcap,
                              const float* dpre, float* dW, float* db, void* workspace, int B, int Vin,
                              int Vout, int S, int Cin, int Cout, sdvae_stream_t stream);

/* Narrow-output SpiralConv forward (Cin = 32, S*Cout <= 27: the 32 -> 3 output layer, model.py:135-136, 172 through
 * model.py:27-41) on tcgen05 by project-then-gather: the tile's distinct staged source rows (FORWARD tile plan of the
 * layer's table, tables.tile_plan, rcap <= 256) are multiplied by the 32 x (S*Cout) weight block once, every output
 * row sums its S Cout-vectors from shared memory.  W is the layer's own [Cout, S*32] weight (no packed image), no
 * activation (the output layer has none).  y [B, Vout, Cout]. */
int sdvae_narrow_out_fwd_tc_supported(int S, int Cin, int Cout, int rcap);
int sdvae_narrow_out_fwd_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                            const uint32_t* plan_cell, int rcap, const float* W, const float* bias, float* y,
                            int B, int Vin, int Vout, int S, int Cin, int Cout, sdvae_stream_t stream);

/* Narrow-output SpiralConv backward (Cin = 32, Cout <= 3, S <= 9: autograd of model.py:34,40 for the 32 -> 3 output
 * layer) on tcgen05 in one fused pass: G[u, s*Cout+n] = in-order sum of dy over the inverse cells (INVERSE tile plan
 * of the layer's table: tables.tile_plan of inverse_cells, rcap <= 288), dx = (G W') * elu'(x) when `gate` (x is the ELU
 * output that fed the layer), dW = sum_u G x, db = sum dy.  dy [B, Vrows, Cout], x and dx [B, Vdst, 32], dW [Cout, S*32],
 * db [Cout]; workspace: sdvae_narrow_out_bwd_tc_workspace bytes.  Deterministic. */
int sdvae_narrow_out_bwd_tc_supported(int S, int Cin, int Cout, int rcap, int ecap);
size_t sdvae_narrow_out_bwd_tc_workspace(int S, int Cout);
int sdvae_narrow_out_bwd_tc(const float* dy, const float* x, const int32_t* plan_cnt, const int32_t* plan_src,
                            const uint32_t* plan_cell, const uint16_t* plan_ext, int rcap, int ecap, const float* W,
                            float* dx, float* dW, float* db, void* workspace, int B, int Vrows, int Vdst, int S,
                            int Cin, int Cout, int gate, sdvae_stream_t stream);

/* ---- narrow-channel layers (C = 3: model.py:104-106 first encoder block, model.py:135-136 output
 * layer) through slot packing: the S*C <= 32 gathered columns of a vertex are materialised once as a
 * 128-byte row, after which every pass of the layer is a dense 32 x 32 contraction on the tcgen05
 * kernels with an identity (S = 1) tile plan.
 *   slot_pack   out[b, r, s*C + c] = sum_{e in cell(r,s)} in[b, cell_src[e], c], columns S*C..31 zero,
 *               out is [B, R, 32]; cell_ptr == NULL: forward table, cell_src = idx[R, S]
 *               (replaces the index_select of model.py:34 for this layer; bit-exact: storage-order sums)
 *   slot_weight Wd[32,32] from the layer weight: mode 0 (narrow input, W [N, S*C]): zero-padded copy;
 *               mode 1 (narrow output, W [C, S*32]): Wd[c, s*C + n] = W[n, s*32 + c]
 *   slot_grad   the inverse scatter of a dense [32,32] weight gradient (+ bias gradient) into dW / db
 *   dense_tc    y[b, r, :] = epi(x[b, r, :32] Wd^T): wimg = sdvae_tc_pack_weights(Wd, S=1, 32, 32, 0),
 *               plan = forward plan of the identity table (S = 1, R rows); gate != NULL: y *= elu'(gate)
 *               (no bias), else bias + optional ELU. */
int sdvae_slot_pack(const float* in, const int32_t* cell_ptr, const int32_t* cell_src, float* out, int B,
                    int Vin, int R, int S, int C, sdvae_stream_t stream);
int sdvae_slot_weight(const float* W, float* Wd, int mode, int N, int S, int C, sdvae_stream_t stream);
int sdvae_slot_grad(const float* dWd, const float* dbd, float* dW, float* db, int mode, int N, int S, int C,
                    sdvae_stream_t stream);
int sdvae_dense_tc(const float* x, const int32_t* plan_cnt, const int32_t* plan_src, int rcap,
                   const float* wimg, const float* bias, const float* gate, float* y, int B, int R,
                   int act, sdvae_stream_t stream);

/* dW[o, s*Cin+c] = sum_{b,v} dpre[b,v,o] * x[b, idx[v,s], c];  db[o] = sum_{b,v} dpre[b,v,o]
 * workspace: sdvae_spiralconv_bwd_w_workspace(...) bytes.  Split-M partial sums are added in a
 * fixed order.  db may be NULL.
 * Replaces: autograd of model.py:40 (grad_weight / grad_bias GEMMs over the materialised gather). */
size_t sdvae_spiralconv_bwd_w_workspace(long long M, int S, int Cin, int Cout);
int sdvae_spiralconv_bwd_w(const float* x, const int32_t* idx, const float* dpre, float* dW,
                           float* db, void* workspace, int B, int Vin, int Vout, int S, int Cin,
                           int Cout, sdvae_stream_t stream);

/* out[m, n] = act(bias[n] + sum_k in[m,k] * W[n*ldw + k]),  dense [M,K] x [K,N] on the same
 * tiled kernel (identity gather).  Used for the per-slot input gradients of the fused encoder
 * blocks.  K must be 32 or 64 for the tiled path (anything else takes the generic kernel). */
int sdvae_dense_fwd(const float* in, const float* W, const float* bias, float* out, long long M,
                    int K, int N, int ldw, int act, sdvae_stream_t stream);

/* out[c, r] = in[r, c] */
int sdvae_transpose2d(const float* in, float* out, int R, int C, sdvae_stream_t stream);

/* ---- narrow-output layer (32 -> 3, model.py:135-136): the whole backward in one pass --------------
 * G[u, s*Cout + n] = sum_{v in cell(u,s)} dy[b, v, n]   (cell_ptr [Vin*S+1], cell_src: the inverse spiral table
 *                                                         in cell-CSR form, rows ascending inside a cell)
 * dx[b,u,c] = (gated ? elu'(x[b,u,c]) : 1) * sum_j G[u,j] * W[j % Cout, (j / Cout)*32 + c]
 * dW[n, s*32+c] = sum_{b,u} G[u, s*Cout+n] * x[b,u,c];   db[n] = sum_{b,v} dy[b,v,n]
 * cell_pack [Vin*S, 2] int32: the first four rows of every cell as 16-bit fields (low half first) holding the
 *   element offset row*Cout; R*Cout = no row; 0xFFFF in the fourth field = the cell has more rows, read from
 *   cell_ptr/cell_src from its 4th on (tables.pack_cells16 builds it; needs R*Cout < 0xFFFF).
 * dy [B,R,Cout], x / dx [B,Vin,32] (x = the layer's input; with gated = 1 it is also the ELU output whose
 * derivative gates dx).  dx, dW, db may each be NULL.  workspace: sdvae_narrow_out_bwd_workspace bytes.
 * Deterministic (fixed summation orders).  Supported for S = 9, Cin = 32, Cout = 3 while one mesh of dy fits
 * shared memory (sdvae_narrow_out_bwd_supported); otherwise use the slot-packed or generic entry points.
 * Replaces: autograd of model.py:34,40 for the output layer (index_select backward = index_add_ atomics,
 * grad_weight GEMM over the materialised gather). */
int sdvae_narrow_out_bwd_supported(int R, int S, int Cin, int Cout);
size_t sdvae_narrow_out_bwd_workspace(int S, int Cout);
int sdvae_narrow_out_bwd(const float* dy, const float* x, const int32_t* cell_ptr, const int32_t* cell_src,
                         const int32_t* cell_pack, const float* W, float* dx, float* dW, float* db, void* workspace, int B, int R,
                         int Vin, int S, int Cin, int Cout, int gated, sdvae_stream_t stream);

/* Forward of the same layer, y[b,v,n] = bias[n] + sum_{s,c} W[n, s*32+c] * x[b, idx[v,s], c], with the DISTINCT
 * source rows of every tile of T = sdvae_narrow_out_fwd_tile() output rows staged once per mesh in shared
 * memory (plan: tables.gather_stage_plan on the spiral table: tile_ptr [L+1], stage_src [tile_ptr[L]] ascending
 * source rows per tile, loc [Vout,S] = position of idx[v,s] in its tile's list, ucap = max rows per tile).
 * fp32 FMA, deterministic.  x [B,Vin,32], out [B,Vout,3]; bias may be NULL.
 * Replaces: model.py:27-41 for de_layers[-1] (index_select -> [B, V*S, Cin] gather -> Linear). */
int sdvae_narrow_out_fwd_tile(void);
int sdvae_narrow_out_fwd_supported(int S, int Cin, int Cout, int ucap);
int sdvae_narrow_out_fwd(const float* x, const int32_t* tile_ptr, const int32_t* stage_src, const int32_t* loc,
                         const float* W, const float* bias, float* out, int B, int Vin, int Vout, int S, int Cin,
                         int Cout, int T, int ucap, sdvae_stream_t stream);

/* ---- narrow-input layer (3 -> 32, model.py:104-110): forward and weight gradient -------------------
 * y[b,r,o] = act(bias[o] + sum_{s,c} W[o, s*Cin+c] * x[b, idx[r,s], c]);   idx [R,S] (may be a row-restricted
 * table), x [B,Vin,3], y [B,R,32].  dW[o, s*Cin+c] = sum_{b,r} dpre[b,r,o] * x[b, idx[r,s], c], db[o] = sum dpre.
 * One mesh of x resident in shared memory, fp32 FMA, deterministic; supported (sdvae_narrow_in_supported) for
 * S = 9, Cin = 3, Cout = 32 while a mesh fits.  dW / db may be NULL.  workspace: sdvae_narrow_in_bwd_w_workspace.
 * Replaces: model.py:27-41 + F.elu for en_layers[0], and autograd's grad_weight GEMM over its gather. */
int sdvae_narrow_in_supported(int Vin, int S, int Cin, int Cout);
size_t sdvae_narrow_in_bwd_w_workspace(int S, int Cin);
int sdvae_narrow_in_fwd(const float* x, const int32_t* idx, const float* W, const float* bias, float* y, int B,
                        int Vin, int R, int S, int Cin, int Cout, int act, sdvae_stream_t stream);
int sdvae_narrow_in_bwd_w(const float* x, const int32_t* idx, const float* dpre, float* dW, float* db,
                          void* workspace, int B, int Vin, int R, int S, int Cin, int Cout,
                          sdvae_stream_t stream);

/* ---- Pool --------------------------------------------------------------------------------- */

/* out[b,r,:] = sum_{j<Wd, col[r,j]>=0} val[r,j] * x[b, col[r,j], :]   (entries in storage order,
 * product rounded before each add).  x [B,Vin,C], col/val [Vout,Wd] (ELL, -1 padding).
 * Replaces: model.py:50-55 (index_select * value -> torch_scatter.scatter_add). */
int sdvae_pool_ell_fwd(const float* x, const int32_t* col, const float* val, float* out, int B,
                       int Vin, int Vout, int Wd, int C, sdvae_stream_t stream);

/* The same result (bit-identical: same products, same order of additions) with the DISTINCT source
 * rows of every tile of T = sdvae_pool_stage_tile() consecutive output rows staged once per mesh in
 * shared memory; the plan is built on the host from the ELL rows (tables.pool_stage_plan):
 *   tile_ptr [L+1], L = ceil(Vout/T); stage_src [tile_ptr[L]]: source rows of tile t, ascending;
 *   ent [Vout,Wd,2] int32: {position of the entry's source row in its tile's list or -1, fp32 value bits}
 *   ucap = max rows staged by a tile.  Supported (sdvae_pool_stage_supported) for C in {32, 64},
 *   2 <= Wd <= 4, when the ring of stage buffers fits shared memory; otherwise call sdvae_pool_ell_fwd.
 * Replaces: model.py:50-55, as above. */
int sdvae_pool_stage_tile(void);
int sdvae_pool_stage_supported(int C, int Wd, int ucap);
int sdvae_pool_ell_fwd_staged(const float* x, const int32_t* tile_ptr, const int32_t* stage_src,
                              const int32_t* ent, float* out, int B, int Vin, int Vout, int Wd, int C,
                              int T, int ucap, sdvae_stream_t stream);

/* dx[b,k,:] = gate * sum_{e in [ptr[k],ptr[k+1])} val[e] * dy[b, src[e], :]
 * CSR of the TRANSPOSED matrix, entries in storage order; val NULL = all ones; gate NULL = none.
 * Replaces: autograd of model.py:53-54 (scatter_add / index_select backward, atomics). */
int sdvae_csr_rowsum(const float* dy, const int32_t* ptr, const int32_t* src, const float* val,
                     const float* gate, float* dx, int B, int Vsrc, int Vdst, int C,
                     sdvae_stream_t stream);

/* ---- elementwise ---------------------------------------------------------------------------- */
int sdvae_elu_fwd(const float* x, float* y, long long n, sdvae_stream_t stream);          /* model.py:68,84 */
int sdvae_elu_bwd(const float* dy, const float* y, float* dx, long long n, sdvae_stream_t stream);
/* z = mu + eps * exp(logvar/2)                                           model.py:184-188 */
int sdvae_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z,
                      long long n, sdvae_stream_t stream);
int sdvae_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu,
                      float* dlogvar, long long n, int accumulate, sdvae_stream_t stream);
/* out = a + sb*b + sc*c (any of a, b, c may be NULL) */
int sdvae_axpy3(const float* a, const float* b, float sb, const float* c, float sc, float* out,
                long long n, sdvae_stream_t stream);

/* ---- feature swap ---------------------------------------------------------------------------- */
/* out[(i-i0)*bs + j, v, :] = mask[v] ? x[j,v,:] : x[i,v,:]  for i in [i0,i1)
 * Replaces: swap_batch_transform.py:13-52 (CPU double loop in the DataLoader collate). */
int sdvae_swap(const float* x, const uint8_t* mask, float* out, int bs, int i0, int i1, int V,
               int C, sdvae_stream_t stream);

/* ---- losses (slots follow ModelManager.loss_keys, model_manager.py:150-154:
 *      0 reconstruction, 1 kl, 2 latent_consistency, 3 laplacian, 4 classification,
 *      5 classification_acc, 6 tot) ------------------------------------------------------------ */

/* Replaces: ModelManager._compute_l1_loss (model_manager.py:328-330, torch.nn.L1Loss(reduction='mean')): out[0] =
 * mean |a - b| over n elements (partial: ceil(n / 256) floats of workspace); l1_bwd: da = g * sign(a - b) / n. */
int sdvae_l1_fwd(const float* a, const float* b, float* partial, float* out, long long n, sdvae_stream_t stream);
int sdvae_l1_bwd(const float* a, const float* b, float* da, long long n, float g, sdvae_stream_t stream);

/* losses[0] = mean((recon-x)^2)          model_manager.py:332-334
 * losses[3] = sum_b sum_v |(L recon_b)_v| / V / B   (if lcol != NULL)   model_manager.py:343-349
 * lcol/lval [V,lw] ELL of the random-walk Laplacian; qn [B,V,3] receives q/|q| for the backward;
 * partial: sdvae_mse_lap_partial_floats(B,V) floats. */
size_t sdvae_mse_lap_partial_floats(int B, int V);
int sdvae_mse_lap_fwd(const float* recon, const float* x, const int32_t* lcol, const float* lval,
                      int lw, float* qn, float* partial, float* losses, int B, int V,
                      float inv_count_scale, sdvae_stream_t stream);
/* drecon = g_mse * d mse/d recon + g_lap * d lap/d recon; tptr/trow/tval = CSR of L^T;
 * dscale (device, 2 floats, may be NULL) multiplies g_mse, g_lap. */
int sdvae_mse_lap_bwd(const float* recon, const float* x, const float* qn, const int32_t* tptr,
                      const int32_t* trow, const float* tval, float* drecon, int B, int V,
                      float g_mse, float g_lap, float inv_count_scale, const float* dscale,
                      sdvae_stream_t stream);

/* losses[1] = mean_b(-1/2 sum_d(1 + lv - mu^2 - exp(lv))); dmu, dlv = its gradients.
 * model_manager.py:351-354.  partial: ceil(B*D/256) floats. */
int sdvae_kl_fwd_bwd(const float* mu, const float* logvar, float* dmu, float* dlogvar,
                     float* partial, float* losses, int B, int D, float inv_count_scale,
                     sdvae_stream_t stream);

/* losses[2] = latent-consistency loss of z [bs*bs, D] on the swap grid, region columns [r0,r1);
 * dz = its gradient.  act_ws: bs*bs*(bs-1) bytes; partial: ceil(bs*bs*(bs-1)/2/256) floats.
 * model_manager.py:360-393. */
int sdvae_lc_fwd_bwd(const float* z, int bs, int D, int r0, int r1, float eta1, float eta2,
                     uint8_t* act_ws, float* partial, float* dz, float* losses,
                     sdvae_stream_t stream);

/* losses[6] = losses[0] + w_kl*losses[1] + w_lc*losses[2] + w_lap*losses[3] + w_cls*losses[4]
 * model_manager.py:308-312 */
int sdvae_total_loss(float* losses, float w_kl, float w_lc, float w_lap, float w_cls,
                     sdvae_stream_t stream);

/* ---- optimiser --------------------------------------------------------------------------------- */
/* torch.optim.Adam step (model_manager.py:69-72, 316) over a flat parameter arena.
 * step_dev: device int holding t (>=1), or NULL to use step_host.  gscale multiplies the gradient
 * (1/world_size after a SUM all-reduce). */
int sdvae_adam_tick(int32_t* step_dev, sdvae_stream_t stream);
int sdvae_adam_step(float* p, const float* grad, float* m, float* v, long long n,
                    const int32_t* step_dev, int step_host, float lr, float beta1, float beta2,
                    float eps, float weight_decay, float gscale, sdvae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SDVAE_B200_H */
