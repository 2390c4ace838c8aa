"""ctypes binding of the C ABI in ``include/sdvae_b200.h``.

This is the only place that touches ``libsdvae_b200.so``.  Every wrapper takes
torch CUDA tensors, checks dtype / contiguity / device, and passes raw device
pointers plus the current CUDA stream.  A non-zero return code becomes a
``RuntimeError`` carrying ``sdvae_last_error()``.  There is no CPU fallback: a
missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdvae_b200.so")

ACT_NONE, ACT_ELU = 0, 1

_lib = None

_c_fp = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "sdvae_abi_version": (C.c_int, []),
    "sdvae_last_error": (C.c_char_p, []),
    "sdvae_spiralconv_fwd": (C.c_int, [_c_fp] * 5 + [C.c_int] * 7 + [_c_fp]),
    "sdvae_weight_transpose": (C.c_int, [_c_fp, _c_fp, C.c_int, C.c_int, C.c_int, _c_fp]),
    "sdvae_spiralconv_bwd_x": (C.c_int, [_c_fp] * 6 + [C.c_int] * 6 + [_c_fp]),
    "sdvae_tc_supported": (C.c_int, [C.c_int] * 4),
    "sdvae_tc_wimg_floats": (C.c_size_t, [C.c_int] * 3),
    "sdvae_tc_pack_weights": (C.c_int, [_c_fp, _c_fp, C.c_int, C.c_int, C.c_int, C.c_int, _c_fp]),
    "sdvae_tc_plan_tiles": (C.c_int, [C.c_int]),
    "sdvae_tc_plan_max_rows": (C.c_int, [_c_fp, C.c_int, C.c_int]),
    "sdvae_tc_plan_build": (C.c_int, [_c_fp, _c_fp, C.c_int, C.c_int, C.c_int, _c_fp, _c_fp, _c_fp]),
    "sdvae_tc_pack_weights_batch": (C.c_int, [_c_fp, C.c_int, _c_fp]),
    "sdvae_tc_pack_weights_part": (C.c_int, [_c_fp, _c_fp] + [C.c_int] * 6 + [_c_fp]),
    "sdvae_spiralconv_fwd_tc": (C.c_int, [_c_fp] * 4 + [C.c_int] + [_c_fp] * 3 + [C.c_int] * 8 + [_c_fp]),
    "sdvae_spiralconv_bwd_x_tc": (C.c_int, [_c_fp] * 4 + [C.c_int] + [_c_fp] * 3 + [C.c_int] * 7 + [_c_fp]),
    "sdvae_tc_bwd_w_supported": (C.c_int, [C.c_int] * 4),
    "sdvae_spiralconv_bwd_w_tc": (C.c_int, [_c_fp] * 3 + [C.c_int] + [_c_fp] * 4 + [C.c_int] * 6 + [_c_fp]),
    "sdvae_dense_tc": (C.c_int, [_c_fp] * 3 + [C.c_int] + [_c_fp] * 4 + [C.c_int] * 3 + [_c_fp]),
    "sdvae_slot_pack": (C.c_int, [_c_fp] * 4 + [C.c_int] * 5 + [_c_fp]),
    "sdvae_slot_weight": (C.c_int, [_c_fp, _c_fp] + [C.c_int] * 4 + [_c_fp]),
    "sdvae_slot_grad": (C.c_int, [_c_fp] * 4 + [C.c_int] * 4 + [_c_fp]),
    "sdvae_spiralconv_bwd_w_workspace": (C.c_size_t, [C.c_longlong, C.c_int, C.c_int, C.c_int]),
    "sdvae_spiralconv_bwd_w": (C.c_int, [_c_fp] * 6 + [C.c_int] * 6 + [_c_fp]),
    "sdvae_dense_fwd": (C.c_int, [_c_fp] * 4 + [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, _c_fp]),
    "sdvae_transpose2d": (C.c_int, [_c_fp, _c_fp, C.c_int, C.c_int, _c_fp]),
    "sdvae_pool_ell_fwd": (C.c_int, [_c_fp] * 4 + [C.c_int] * 5 + [_c_fp]),
    "sdvae_narrow_out_bwd_supported": (C.c_int, [C.c_int] * 4),
    "sdvae_narrow_out_bwd_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "sdvae_narrow_out_bwd": (C.c_int, [_c_fp] * 10 + [C.c_int] * 7 + [_c_fp]),
    "sdvae_tile_supported": (C.c_int, [C.c_int] * 5),
    "sdvae_tile_bwd_w_supported": (C.c_int, [C.c_int] * 4),
    "sdvae_spiralconv_bwd_w_tile": (C.c_int, [_c_fp] * 4 + [C.c_int] + [_c_fp] * 4 + [C.c_int] * 6 + [_c_fp]),
    "sdvae_spiralconv_fwd_tile": (C.c_int, [_c_fp] * 4 + [C.c_int] + [_c_fp] * 3 + [C.c_int] * 7 + [_c_fp]),
    "sdvae_spiralconv_bwd_x_tile": (C.c_int, [_c_fp] * 5 + [C.c_int] * 2 + [_c_fp] * 3 + [C.c_int] * 6 + [_c_fp]),
    "sdvae_narrow_out_fwd_tc_supported": (C.c_int, [C.c_int] * 4),
    "sdvae_narrow_out_fwd_tc": (C.c_int, [_c_fp] * 4 + [C.c_int] + [_c_fp] * 3 + [C.c_int] * 6 + [_c_fp]),
    "sdvae_narrow_out_bwd_tc_supported": (C.c_int, [C.c_int] * 5),
    "sdvae_narrow_out_bwd_tc_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "sdvae_narrow_out_bwd_tc": (C.c_int, [_c_fp] * 6 + [C.c_int] * 2 + [_c_fp] * 5 + [C.c_int] * 7 + [_c_fp]),
    "sdvae_narrow_in_supported": (C.c_int, [C.c_int] * 4),
    "sdvae_narrow_in_bwd_w_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "sdvae_narrow_in_fwd": (C.c_int, [_c_fp] * 5 + [C.c_int] * 7 + [_c_fp]),
    "sdvae_narrow_in_bwd_w": (C.c_int, [_c_fp] * 6 + [C.c_int] * 6 + [_c_fp]),
    "sdvae_narrow_out_fwd_tile": (C.c_int, []),
    "sdvae_narrow_out_fwd_supported": (C.c_int, [C.c_int] * 4),
    "sdvae_narrow_out_fwd": (C.c_int, [_c_fp] * 7 + [C.c_int] * 8 + [_c_fp]),
    "sdvae_pool_stage_tile": (C.c_int, []),
    "sdvae_pool_stage_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "sdvae_pool_ell_fwd_staged": (C.c_int, [_c_fp] * 5 + [C.c_int] * 7 + [_c_fp]),
    "sdvae_csr_rowsum": (C.c_int, [_c_fp] * 6 + [C.c_int] * 4 + [_c_fp]),
    "sdvae_elu_fwd": (C.c_int, [_c_fp, _c_fp, C.c_longlong, _c_fp]),
    "sdvae_elu_bwd": (C.c_int, [_c_fp, _c_fp, _c_fp, C.c_longlong, _c_fp]),
    "sdvae_reparam_fwd": (C.c_int, [_c_fp] * 4 + [C.c_longlong, _c_fp]),
    "sdvae_reparam_bwd": (C.c_int, [_c_fp] * 5 + [C.c_longlong, C.c_int, _c_fp]),
    "sdvae_axpy3": (C.c_int, [_c_fp, _c_fp, C.c_float, _c_fp, C.c_float, _c_fp, C.c_longlong, _c_fp]),
    "sdvae_swap": (C.c_int, [_c_fp] * 3 + [C.c_int] * 5 + [_c_fp]),
    "sdvae_l1_fwd": (C.c_int, [_c_fp] * 4 + [C.c_longlong, _c_fp]),
    "sdvae_l1_bwd": (C.c_int, [_c_fp] * 3 + [C.c_longlong, C.c_float, _c_fp]),
    "sdvae_mse_lap_partial_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "sdvae_mse_lap_fwd": (C.c_int, [_c_fp] * 4 + [C.c_int] + [_c_fp] * 3 + [C.c_int, C.c_int, C.c_float, _c_fp]),
    "sdvae_mse_lap_bwd": (C.c_int, [_c_fp] * 7 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _c_fp, _c_fp]),
    "sdvae_kl_fwd_bwd": (C.c_int, [_c_fp] * 6 + [C.c_int, C.c_int, C.c_float, _c_fp]),
    "sdvae_lc_fwd_bwd": (C.c_int, [_c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float] + [_c_fp] * 4 + [_c_fp]),
    "sdvae_total_loss": (C.c_int, [_c_fp, C.c_float, C.c_float, C.c_float, C.c_float, _c_fp]),
    "sdvae_adam_tick": (C.c_int, [_c_fp, _c_fp]),
    "sdvae_adam_step": (C.c_int, [_c_fp] * 4 + [C.c_longlong, _c_fp, C.c_int] + [C.c_float] * 6 + [_c_fp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load(build_if_missing: bool = False):
    """Load the shared library (once).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build_library()
        else:
            raise RuntimeError(
                "sdvae_b200: %s is missing -- run `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# kernels launched per ABI call (for bench.py's `gpu_launches`)
_KERNELS_PER_CALL = {
    "spiralconv_fwd": 1, "weight_transpose": 1, "spiralconv_bwd_x": 1, "spiralconv_bwd_w": 2,
    "tc_pack_weights": 1, "spiralconv_fwd_tc": 1, "spiralconv_bwd_x_tc": 1,
    "spiralconv_bwd_w_tc": 2, "dense_tc": 1, "slot_pack": 1, "slot_weight": 1, "slot_grad": 1,
    "dense_fwd": 1, "transpose2d": 1, "spiralconv_fwd_tile": 1, "spiralconv_bwd_x_tile": 1, "spiralconv_bwd_w_tile": 2, "narrow_out_bwd": 2, "narrow_out_fwd": 1, "narrow_in_fwd": 1, "narrow_in_bwd_w": 2, "pool_ell_fwd": 1, "pool_ell_fwd_staged": 1, "csr_rowsum": 1, "elu_fwd": 1,
    "elu_bwd": 1, "reparam_fwd": 1, "reparam_bwd": 1, "axpy3": 1, "swap": 1, "mse_lap_fwd": 3,
    "mse_lap_bwd": 1, "kl_fwd_bwd": 2, "lc_fwd_bwd": 3, "total_loss": 1, "adam_tick": 1,
    "adam_step": 1,
}
_launches = 0


def launch_count() -> int:
    """Number of sdvae_b200 kernels launched (or replayed from a graph) by this process."""
    return _launches


def add_launches(n: int):
    global _launches
    _launches += int(n)


def _err(code: int, what: str):
    msg = load().sdvae_last_error().decode("utf-8", "replace")
    raise RuntimeError("sdvae_b200 %s failed (code %d): %s" % (what, code, msg))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str) -> int:
    if not t.is_cuda:
        raise RuntimeError("sdvae_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
    if t.device.index != torch.cuda.current_device():
        # the launch goes to the current device's current stream: a tensor of another device would be read through a
        # foreign pointer.  TrainEngine / the autograd functions set the device of their tensors before they call.
        raise RuntimeError("sdvae_b200: %s is on %s but the current device is cuda:%d (use torch.cuda.device(...))"
                           % (name, t.device, torch.cuda.current_device()))
    if t.dtype != dtype:
        raise TypeError("sdvae_b200: %s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("sdvae_b200: %s must be contiguous" % name)
    return t.data_ptr()


def _f(t, name):
    return _chk(t, torch.float32, name)


def _i(t, name):
    return _chk(t, torch.int32, name)


def _fo(t: Optional[torch.Tensor], name):
    return None if t is None else _f(t, name)


def _io(t: Optional[torch.Tensor], name):
    return None if t is None else _i(t, name)


# ---------------------------------------------------------------------------------
def spiralconv_fwd(x, idx32, weight, bias, y, B, Vin, Vout, S, Cin, Cout, act):
    rc = load().sdvae_spiralconv_fwd(_f(x, "x"), _i(idx32, "idx"), _f(weight, "weight"),
                                     _fo(bias, "bias"), _f(y, "y"), B, Vin, Vout, S, Cin, Cout,
                                     act, _stream())
    if rc:
        _err(rc, "spiralconv_fwd")
    add_launches(_KERNELS_PER_CALL["spiralconv_fwd"])


def weight_transpose(weight, wt, Cout, Cin, S):
    rc = load().sdvae_weight_transpose(_f(weight, "weight"), _f(wt, "wt"), Cout, Cin, S, _stream())
    if rc:
        _err(rc, "weight_transpose")
    add_launches(_KERNELS_PER_CALL["weight_transpose"])


def spiralconv_bwd_x(dpre, cell_ptr, cell_src, wt, gate, dx, B, Vrows, Vdst, S, Cout, Cin):
    rc = load().sdvae_spiralconv_bwd_x(_f(dpre, "dpre"), _i(cell_ptr, "cell_ptr"),
                                       _i(cell_src, "cell_src"), _f(wt, "wt"), _fo(gate, "gate"),
                                       _f(dx, "dx"), B, Vrows, Vdst, S, Cout, Cin, _stream())
    if rc:
        _err(rc, "spiralconv_bwd_x")
    add_launches(_KERNELS_PER_CALL["spiralconv_bwd_x"])


# ---- tensor-core (tcgen05) SpiralConv ---------------------------------------------
def tc_supported(S, KS, N, rcap=128) -> bool:
    return bool(load().sdvae_tc_supported(S, KS, N, rcap))


def tc_wimg_floats(S, KS, N) -> int:
    return int(load().sdvae_tc_wimg_floats(S, KS, N))


def tc_pack_weights(weight, wimg, S, Cin, Cout, transposed, n0=0, n_cnt=None, kperm=False):
    """Packed weight image of output channels [n0, n0 + n_cnt) (input channels if ``transposed``);
    ``kperm``: the K order of the tile-staged kernels (csrc/spiral_conv_tile.cuh)."""
    if n_cnt is None:
        n_cnt = (Cin if transposed else Cout) - n0
    rc = load().sdvae_tc_pack_weights_part(_f(weight, "weight"), _f(wimg, "wimg"), S, Cin, Cout,
                                           (1 if transposed else 0) | (2 if kperm else 0), n0, n_cnt, _stream())
    if rc:
        _err(rc, "tc_pack_weights")
    add_launches(_KERNELS_PER_CALL["tc_pack_weights"])


class PackEntry(C.Structure):
    """``sdvae_pack_entry`` of include/sdvae_b200.h."""
    _fields_ = [("W", C.c_void_p), ("wimg", C.c_void_p), ("S", C.c_int), ("Cin", C.c_int),
                ("Cout", C.c_int), ("transposed", C.c_int), ("n0", C.c_int), ("n_cnt", C.c_int)]


def tc_pack_table(entries, device) -> torch.Tensor:
    """Device table for ``tc_pack_weights_batch`` from ``(weight, wimg, S, Cin, Cout, transposed, n0,
    n_cnt)`` tuples.  The tensors must outlive the table (it holds raw pointers)."""
    arr = (PackEntry * len(entries))()
    for i, (w, img, S, Cin, Cout, tr, n0, nc) in enumerate(entries):
        # tr: bool (transposed) or the flag word of sdvae_pack_entry (bit 0 transposed, bit 1 kperm image)
        arr[i] = PackEntry(_f(w, "weight"), _f(img, "wimg"), S, Cin, Cout, int(tr), n0, nc)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8) if len(entries) else torch.zeros(0, dtype=torch.uint8)
    return host.to(device)


def tc_pack_weights_batch(table: torch.Tensor, n: int):
    if n == 0:
        return
    rc = load().sdvae_tc_pack_weights_batch(_chk(table, torch.uint8, "table"), n, _stream())
    if rc:
        _err(rc, "tc_pack_weights_batch")
    add_launches(1)


def tc_plan_build(cell_ptr, cell_src, out_rows, S):
    """Host-side tile plan of a cell-form table (NumPy int32 arrays).  Returns
    ``(cnt [L,S], src [L,S,rcap/2] (packed), cell [L,S,128], rcap)`` as NumPy arrays."""
    import numpy as np
    lib = load()
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int32)
    cell_src = np.ascontiguousarray(cell_src, np.int32)
    if cell_ptr.shape[0] != out_rows * S + 1:
        raise ValueError("tc_plan_build: cell_ptr must have out_rows*S + 1 entries")
    L = int(lib.sdvae_tc_plan_tiles(out_rows))
    mx = int(lib.sdvae_tc_plan_max_rows(cell_ptr.ctypes.data, out_rows, S))
    if mx < 0:
        raise RuntimeError("tc_plan_build: bad table")
    rcap = max(32, (mx + 31) // 32 * 32)
    cnt = np.zeros((L, S), np.int32)
    src = np.zeros((L, S, rcap // 2), np.int32)       # packed, two 16-bit rows per word
    cell = np.zeros((L, S, 128), np.int32)
    rc = lib.sdvae_tc_plan_build(cell_ptr.ctypes.data, cell_src.ctypes.data, out_rows, S, rcap,
                                 cnt.ctypes.data, src.ctypes.data, cell.ctypes.data)
    if rc:
        _err(rc, "tc_plan_build")
    return cnt, src, cell, rcap


def spiralconv_fwd_tc(x, plan, wimg, bias, y, B, Vin, Vout, S, Cin, Cout, act, ldy=0):
    """``ldy`` > Cout: ``y`` (and ``bias``) start at the first of Cout consecutive output channels of a wider
    [.., ldy] tensor (a flat view sliced at that channel)."""
    rc = load().sdvae_spiralconv_fwd_tc(_f(x, "x"), _i(plan.cnt, "plan.cnt"), _i(plan.src, "plan.src"),
                                        _i(plan.cell, "plan.cell"), plan.rcap, _f(wimg, "wimg"),
                                        _fo(bias, "bias"), _f(y, "y"), B, Vin, Vout, S, Cin, Cout,
                                        act, ldy, _stream())
    if rc:
        _err(rc, "spiralconv_fwd_tc")
    add_launches(_KERNELS_PER_CALL["spiralconv_fwd_tc"])


def spiralconv_bwd_x_tc(dpre, plan, wimg_t, gate, dx, B, Vrows, Vdst, S, Cout, Cin, lddx=0):
    rc = load().sdvae_spiralconv_bwd_x_tc(_f(dpre, "dpre"), _i(plan.cnt, "plan.cnt"),
                                          _i(plan.src, "plan.src"), _i(plan.cell, "plan.cell"),
                                          plan.rcap, _f(wimg_t, "wimg_t"), _fo(gate, "gate"),
                                          _f(dx, "dx"), B, Vrows, Vdst, S, Cout, Cin, lddx, _stream())
    if rc:
        _err(rc, "spiralconv_bwd_x_tc")
    add_launches(_KERNELS_PER_CALL["spiralconv_bwd_x_tc"])


def tc_bwd_w_supported(S, Cin, Cout, rcap=128) -> bool:
    return bool(load().sdvae_tc_bwd_w_supported(S, Cin, Cout, rcap))


def spiralconv_bwd_w_tc(x, plan, dpre, dW, db, workspace, B, Vin, Vout, S, Cin, Cout):
    """Weight / bias gradient on the tcgen05 path; ``plan`` is the layer's FORWARD tile plan."""
    rc = load().sdvae_spiralconv_bwd_w_tc(_f(x, "x"), _i(plan.cnt, "plan.cnt"), _i(plan.src, "plan.src"),
                                          plan.rcap, _f(dpre, "dpre"), _f(dW, "dW"), _fo(db, "db"),
                                          _f(workspace, "workspace"), B, Vin, Vout, S, Cin, Cout,
                                          _stream())
    if rc:
        _err(rc, "spiralconv_bwd_w_tc")
    add_launches((Cin // 32) * ((Cout + 31) // 32) + 1)      # one bw_umma_kernel per 32 x <=32 pass + the reduction


def dense_tc(x, plan, wimg, bias, gate, y, B, R, act):
    """y[b, r, :] = epi(x[b, r, :32] Wd^T) on the tcgen05 path; ``plan`` = forward plan of the identity
    table with R rows (S = 1)."""
    rc = load().sdvae_dense_tc(_f(x, "x"), _i(plan.cnt, "plan.cnt"), _i(plan.src, "plan.src"), plan.rcap,
                               _f(wimg, "wimg"), _fo(bias, "bias"), _fo(gate, "gate"), _f(y, "y"), B, R,
                               act, _stream())
    if rc:
        _err(rc, "dense_tc")
    add_launches(_KERNELS_PER_CALL["dense_tc"])


def slot_pack(inp, cell_ptr, cell_src, out, B, Vin, R, S, Cn):
    rc = load().sdvae_slot_pack(_f(inp, "in"), _io(cell_ptr, "cell_ptr"), _i(cell_src, "cell_src"),
                                _f(out, "out"), B, Vin, R, S, Cn, _stream())
    if rc:
        _err(rc, "slot_pack")
    add_launches(_KERNELS_PER_CALL["slot_pack"])


def slot_weight(W, Wd, mode, N, S, Cn):
    rc = load().sdvae_slot_weight(_f(W, "W"), _f(Wd, "Wd"), mode, N, S, Cn, _stream())
    if rc:
        _err(rc, "slot_weight")
    add_launches(_KERNELS_PER_CALL["slot_weight"])


def slot_grad(dWd, dbd, dW, db, mode, N, S, Cn):
    rc = load().sdvae_slot_grad(_f(dWd, "dWd"), _fo(dbd, "dbd"), _f(dW, "dW"), _fo(db, "db"), mode, N, S,
                                Cn, _stream())
    if rc:
        _err(rc, "slot_grad")
    add_launches(_KERNELS_PER_CALL["slot_grad"])


def spiralconv_bwd_w_workspace(M, S, Cin, Cout) -> int:
    return int(load().sdvae_spiralconv_bwd_w_workspace(M, S, Cin, Cout))


def spiralconv_bwd_w(x, idx32, dpre, dW, db, workspace, B, Vin, Vout, S, Cin, Cout):
    rc = load().sdvae_spiralconv_bwd_w(_f(x, "x"), _i(idx32, "idx"), _f(dpre, "dpre"),
                                       _f(dW, "dW"), _fo(db, "db"), _f(workspace, "workspace"),
                                       B, Vin, Vout, S, Cin, Cout, _stream())
    if rc:
        _err(rc, "spiralconv_bwd_w")
    add_launches(_KERNELS_PER_CALL["spiralconv_bwd_w"])


def dense_fwd(inp, weight, bias, out, M, K, N, ldw, act):
    rc = load().sdvae_dense_fwd(_f(inp, "in"), _f(weight, "weight"), _fo(bias, "bias"),
                                _f(out, "out"), M, K, N, ldw, act, _stream())
    if rc:
        _err(rc, "dense_fwd")
    add_launches(_KERNELS_PER_CALL["dense_fwd"])


def transpose2d(inp, out, R, Ccols):
    rc = load().sdvae_transpose2d(_f(inp, "in"), _f(out, "out"), R, Ccols, _stream())
    if rc:
        _err(rc, "transpose2d")
    add_launches(_KERNELS_PER_CALL["transpose2d"])


def pool_ell_fwd(x, col, val, out, B, Vin, Vout, Wd, Cc):
    rc = load().sdvae_pool_ell_fwd(_f(x, "x"), _i(col, "col"), _f(val, "val"), _f(out, "out"),
                                   B, Vin, Vout, Wd, Cc, _stream())
    if rc:
        _err(rc, "pool_ell_fwd")
    add_launches(_KERNELS_PER_CALL["pool_ell_fwd"])


def tile_supported(S: int, Cin: int, Cout: int, rcap: int, ecap: int = 0) -> bool:
    return bool(load().sdvae_tile_supported(int(S), int(Cin), int(Cout), int(rcap), int(ecap)))


def _i16(t, name):
    return _chk(t, torch.int16, name)


def spiralconv_fwd_tile(x, plan, wimg, bias, y, B, Vin, Vout, S, Cin, Cout, act):
    """SpiralConv (+ELU) forward on tcgen05 with tile-local staging; ``plan`` = ``tables.TileStagePlan`` of the
    forward gather, ``wimg`` packed with ``tc_pack_weights(..., kperm=True)``."""
    rc = load().sdvae_spiralconv_fwd_tile(_f(x, "x"), _i(plan.cnt, "plan_cnt"), _i(plan.src, "plan_src"),
                                          _i(plan.cell, "plan_cell"), plan.rcap, _f(wimg, "wimg"), _fo(bias, "bias"),
                                          _f(y, "y"), B, Vin, Vout, S, Cin, Cout, act, _stream())
    if rc:
        _err(rc, "spiralconv_fwd_tile")
    add_launches(_KERNELS_PER_CALL["spiralconv_fwd_tile"])


def narrow_out_fwd_tc_supported(S, Cin, Cout, rcap):
    return bool(load().sdvae_narrow_out_fwd_tc_supported(S, Cin, Cout, rcap))


def narrow_out_fwd_tc(x, plan, w, bias, y, B, Vin, Vout, S, Cin, Cout):
    """Narrow-output SpiralConv forward (32 -> 3) on tcgen05, project-then-gather; ``plan`` = the forward
    ``tables.TileStagePlan`` of the layer's table, ``w`` the layer's own [Cout, S*32] weight."""
    rc = load().sdvae_narrow_out_fwd_tc(_f(x, "x"), _i(plan.cnt, "plan_cnt"), _i(plan.src, "plan_src"),
                                        _i(plan.cell, "plan_cell"), plan.rcap, _f(w, "w"), _fo(bias, "bias"),
                                        _f(y, "y"), B, Vin, Vout, S, Cin, Cout, _stream())
    if rc:
        _err(rc, "narrow_out_fwd_tc")
    add_launches(1)


def narrow_out_bwd_tc_supported(S, Cin, Cout, rcap, ecap):
    return bool(load().sdvae_narrow_out_bwd_tc_supported(S, Cin, Cout, rcap, ecap))


def narrow_out_bwd_tc_workspace(S, Cout):
    return int(load().sdvae_narrow_out_bwd_tc_workspace(S, Cout))


def narrow_out_bwd_tc(dy, x, plan, w, dx, dW, db, ws, B, Vrows, Vdst, S, Cin, Cout, gate):
    """Narrow-output SpiralConv backward (32 -> 3) on tcgen05 in one pass (dx with the previous ELU', dW, db);
    ``plan`` = the INVERSE ``tables.TileStagePlan`` (``SpiralTable.tile_bwd()``)."""
    rc = load().sdvae_narrow_out_bwd_tc(_f(dy, "dy"), _f(x, "x"), _i(plan.cnt, "plan_cnt"), _i(plan.src, "plan_src"),
                                        _i(plan.cell, "plan_cell"), _i16(plan.ext, "plan_ext"), plan.rcap, plan.ecap,
                                        _f(w, "w"), _f(dx, "dx"), _f(dW, "dW"), _f(db, "db"), _f(ws, "workspace"),
                                        B, Vrows, Vdst, S, Cin, Cout, 1 if gate else 0, _stream())
    if rc:
        _err(rc, "narrow_out_bwd_tc")
    add_launches(2)


def spiralconv_bwd_x_tile(dpre, plan, wimg_t, gate, dx, B, Vrows, Vdst, S, Cout, Cin):
    """Backward-to-input of SpiralConv on tcgen05 with tile-local staging; ``plan`` = ``TileStagePlan`` of the
    inverse table, ``wimg_t`` packed with ``tc_pack_weights(..., transposed=True, kperm=True)``."""
    rc = load().sdvae_spiralconv_bwd_x_tile(_f(dpre, "dpre"), _i(plan.cnt, "plan_cnt"), _i(plan.src, "plan_src"),
                                            _i(plan.cell, "plan_cell"), _i16(plan.ext, "plan_ext"), plan.rcap,
                                            plan.ecap, _f(wimg_t, "wimg_t"), _fo(gate, "gate"), _f(dx, "dx"),
                                            B, Vrows, Vdst, S, Cout, Cin, _stream())
    if rc:
        _err(rc, "spiralconv_bwd_x_tile")
    add_launches(_KERNELS_PER_CALL["spiralconv_bwd_x_tile"])


def narrow_in_supported(Vin: int, S: int, Cin: int, Cout: int) -> bool:
    return bool(load().sdvae_narrow_in_supported(int(Vin), int(S), int(Cin), int(Cout)))


def narrow_in_bwd_w_workspace(S: int, Cin: int) -> int:
    return int(load().sdvae_narrow_in_bwd_w_workspace(int(S), int(Cin)))


def narrow_in_fwd(x, idx, W, bias, y, B, Vin, R, S, Cin, Cout, act):
    """3 -> 32 SpiralConv forward (+ ELU) with the mesh's input resident in shared memory."""
    rc = load().sdvae_narrow_in_fwd(_f(x, "x"), _i(idx, "idx"), _f(W, "W"), _fo(bias, "bias"), _f(y, "y"),
                                    B, Vin, R, S, Cin, Cout, act, _stream())
    if rc:
        _err(rc, "narrow_in_fwd")
    add_launches(_KERNELS_PER_CALL["narrow_in_fwd"])


def narrow_in_bwd_w(x, idx, dpre, dW, db, ws, B, Vin, R, S, Cin, Cout):
    if ws.numel() * 4 < narrow_in_bwd_w_workspace(S, Cin):
        raise RuntimeError("sdvae_b200: narrow_in_bwd_w workspace too small")
    rc = load().sdvae_narrow_in_bwd_w(_f(x, "x"), _i(idx, "idx"), _f(dpre, "dpre"), _fo(dW, "dW"), _fo(db, "db"),
                                      _f(ws, "ws"), B, Vin, R, S, Cin, Cout, _stream())
    if rc:
        _err(rc, "narrow_in_bwd_w")
    add_launches(_KERNELS_PER_CALL["narrow_in_bwd_w"])


def narrow_out_fwd_supported(S: int, Cin: int, Cout: int, ucap: int) -> bool:
    return bool(load().sdvae_narrow_out_fwd_supported(int(S), int(Cin), int(Cout), int(ucap)))


def narrow_out_fwd(x, plan, W, bias, out, B, Vin, Vout, S, Cin, Cout):
    """32 -> 3 SpiralConv forward (no activation); ``plan``: tables.GatherStagePlan of the spiral table."""
    rc = load().sdvae_narrow_out_fwd(_f(x, "x"), _i(plan.tile_ptr, "tile_ptr"), _i(plan.stage_src, "stage_src"),
                                     _i(plan.loc, "loc"), _f(W, "W"), _fo(bias, "bias"), _f(out, "out"),
                                     B, Vin, Vout, S, Cin, Cout, plan.T, plan.ucap, _stream())
    if rc:
        _err(rc, "narrow_out_fwd")
    add_launches(_KERNELS_PER_CALL["narrow_out_fwd"])


def narrow_out_bwd_supported(R: int, S: int, Cin: int, Cout: int) -> bool:
    return bool(load().sdvae_narrow_out_bwd_supported(int(R), int(S), int(Cin), int(Cout)))


def narrow_out_bwd_workspace(S: int, Cout: int) -> int:
    return int(load().sdvae_narrow_out_bwd_workspace(int(S), int(Cout)))


def narrow_out_bwd(dy, x, cell_ptr, cell_src, cell_pack, W, dx, dW, db, ws, B, R, Vin, S, Cin, Cout, gated):
    """Fused backward of the 32 -> 3 output layer (dx, dW, db; any of them may be None)."""
    if ws.numel() * 4 < narrow_out_bwd_workspace(S, Cout):
        raise RuntimeError("sdvae_b200: narrow_out_bwd workspace too small")
    rc = load().sdvae_narrow_out_bwd(_f(dy, "dy"), _f(x, "x"), _i(cell_ptr, "cell_ptr"), _i(cell_src, "cell_src"),
                                     _i(cell_pack, "cell_pack"), _f(W, "W"), _fo(dx, "dx"), _fo(dW, "dW"), _fo(db, "db"), _f(ws, "ws"),
                                     B, R, Vin, S, Cin, Cout, 1 if gated else 0, _stream())
    if rc:
        _err(rc, "narrow_out_bwd")
    add_launches(_KERNELS_PER_CALL["narrow_out_bwd"])


def pool_stage_supported(Cc: int, Wd: int, ucap: int) -> bool:
    return bool(load().sdvae_pool_stage_supported(int(Cc), int(Wd), int(ucap)))


def tile_bwd_w_supported(S: int, Cin: int, Cout: int, rcap: int) -> bool:
    return bool(load().sdvae_tile_bwd_w_supported(int(S), int(Cin), int(Cout), int(rcap)))


def spiralconv_bwd_w_tile(x, plan, dpre, dW, db, ws, B, Vin, Vout, S, Cin, Cout):
    """SpiralConv weight / bias gradient on tcgen05 with tile-local staging; ``plan`` = ``tables.TileStagePlan`` of
    the FORWARD gather (``SpiralTable.tile_fwd()``)."""
    need = spiralconv_bwd_w_workspace(B * Vout, S, Cin, Cout)
    if ws.numel() * 4 < need:
        raise ValueError("spiralconv_bwd_w_tile: workspace too small (%d < %d bytes)" % (ws.numel() * 4, need))
    rc = load().sdvae_spiralconv_bwd_w_tile(_f(x, "x"), _i(plan.cnt, "plan_cnt"), _i(plan.src, "plan_src"),
                                            _i(plan.cell, "plan_cell"), plan.rcap, _f(dpre, "dpre"), _f(dW, "dW"),
                                            _fo(db, "db"), _f(ws, "workspace"), B, Vin, Vout, S, Cin, Cout, _stream())
    if rc:
        _err(rc, "spiralconv_bwd_w_tile")
    add_launches((Cin // 32) * ((Cout + 31) // 32) + 1)      # one bt_kernel per 32 x <=32 pass + the reduction


def pool_ell_fwd_staged(x, plan, out, B, Vin, Vout, Wd, Cc):
    """``plan``: tables.PoolStagePlan (tile_ptr, stage_src, ent, T, ucap)."""
    rc = load().sdvae_pool_ell_fwd_staged(_f(x, "x"), _i(plan.tile_ptr, "tile_ptr"),
                                          _i(plan.stage_src, "stage_src"), _i(plan.ent, "ent"),
                                          _f(out, "out"), B, Vin, Vout, Wd, Cc, plan.T, plan.ucap,
                                          _stream())
    if rc:
        _err(rc, "pool_ell_fwd_staged")
    add_launches(_KERNELS_PER_CALL["pool_ell_fwd_staged"])


POOL_STAGE_MIN_MESHES = 8       # below this the ring of staged meshes never fills; the L2 gather is as fast


def pool_fwd(x, table, out, B, Vin, Cc):
    """Pool forward through the staged kernel when the table has a usable stage plan, else the ELL gather."""
    aligned = x.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0
    plan = table.stage_plan() if (B >= POOL_STAGE_MIN_MESHES and aligned) else None
    if plan is not None and pool_stage_supported(Cc, table.width, plan.ucap):
        pool_ell_fwd_staged(x, plan, out, B, Vin, table.n_rows, table.width, Cc)
    else:
        pool_ell_fwd(x, table.ell_col, table.ell_val, out, B, Vin, table.n_rows, table.width, Cc)


def csr_rowsum(dy, ptr, src, val, gate, dx, B, Vsrc, Vdst, Cc):
    rc = load().sdvae_csr_rowsum(_f(dy, "dy"), _i(ptr, "ptr"), _i(src, "src"), _fo(val, "val"),
                                 _fo(gate, "gate"), _f(dx, "dx"), B, Vsrc, Vdst, Cc, _stream())
    if rc:
        _err(rc, "csr_rowsum")
    add_launches(_KERNELS_PER_CALL["csr_rowsum"])


def elu_fwd(x, y):
    rc = load().sdvae_elu_fwd(_f(x, "x"), _f(y, "y"), x.numel(), _stream())
    if rc:
        _err(rc, "elu_fwd")
    add_launches(_KERNELS_PER_CALL["elu_fwd"])


def elu_bwd(dy, y, dx):
    rc = load().sdvae_elu_bwd(_f(dy, "dy"), _f(y, "y"), _f(dx, "dx"), dy.numel(), _stream())
    if rc:
        _err(rc, "elu_bwd")
    add_launches(_KERNELS_PER_CALL["elu_bwd"])


def reparam_fwd(mu, logvar, eps, z):
    rc = load().sdvae_reparam_fwd(_f(mu, "mu"), _f(logvar, "logvar"), _f(eps, "eps"), _f(z, "z"),
                                  mu.numel(), _stream())
    if rc:
        _err(rc, "reparam_fwd")
    add_launches(_KERNELS_PER_CALL["reparam_fwd"])


def reparam_bwd(dz, logvar, eps, dmu, dlogvar, accumulate=False):
    rc = load().sdvae_reparam_bwd(_f(dz, "dz"), _f(logvar, "logvar"), _f(eps, "eps"),
                                  _f(dmu, "dmu"), _f(dlogvar, "dlogvar"), dz.numel(),
                                  1 if accumulate else 0, _stream())
    if rc:
        _err(rc, "reparam_bwd")
    add_launches(_KERNELS_PER_CALL["reparam_bwd"])


def axpy3(a, b, sb, c, sc, out):
    rc = load().sdvae_axpy3(_fo(a, "a"), _fo(b, "b"), sb, _fo(c, "c"), sc, _f(out, "out"),
                            out.numel(), _stream())
    if rc:
        _err(rc, "axpy3")
    add_launches(_KERNELS_PER_CALL["axpy3"])


def swap(x, mask_u8, out, bs, i0, i1, V, Cc):
    rc = load().sdvae_swap(_f(x, "x"), _chk(mask_u8, torch.uint8, "mask"), _f(out, "out"),
                           bs, i0, i1, V, Cc, _stream())
    if rc:
        _err(rc, "swap")
    add_launches(_KERNELS_PER_CALL["swap"])


def l1_fwd(a, b, partial, out):
    rc = load().sdvae_l1_fwd(_f(a, "a"), _f(b, "b"), _f(partial, "partial"), _f(out, "out"), a.numel(), _stream())
    if rc:
        _err(rc, "l1_fwd")
    add_launches(2)


def l1_bwd(a, b, da, g):
    rc = load().sdvae_l1_bwd(_f(a, "a"), _f(b, "b"), _f(da, "da"), a.numel(), float(g), _stream())
    if rc:
        _err(rc, "l1_bwd")
    add_launches(1)


def mse_lap_partial_floats(B, V) -> int:
    return int(load().sdvae_mse_lap_partial_floats(B, V))


def mse_lap_fwd(recon, x, lcol, lval, lw, qn, partial, losses, B, V, scale=1.0):
    rc = load().sdvae_mse_lap_fwd(_f(recon, "recon"), _f(x, "x"),
                                  None if lcol is None else _i(lcol, "lcol"), _fo(lval, "lval"), lw,
                                  _fo(qn, "qn"), _f(partial, "partial"), _f(losses, "losses"),
                                  B, V, scale, _stream())
    if rc:
        _err(rc, "mse_lap_fwd")
    add_launches(_KERNELS_PER_CALL["mse_lap_fwd"])


def mse_lap_bwd(recon, x, qn, tptr, trow, tval, drecon, B, V, g_mse, g_lap, scale=1.0, dscale=None):
    rc = load().sdvae_mse_lap_bwd(_f(recon, "recon"), _f(x, "x"), _fo(qn, "qn"),
                                  None if tptr is None else _i(tptr, "tptr"),
                                  None if trow is None else _i(trow, "trow"), _fo(tval, "tval"),
                                  _f(drecon, "drecon"), B, V, g_mse, g_lap, scale,
                                  _fo(dscale, "dscale"), _stream())
    if rc:
        _err(rc, "mse_lap_bwd")
    add_launches(_KERNELS_PER_CALL["mse_lap_bwd"])


def kl_fwd_bwd(mu, logvar, dmu, dlogvar, partial, losses, B, D, scale=1.0):
    rc = load().sdvae_kl_fwd_bwd(_f(mu, "mu"), _f(logvar, "logvar"), _f(dmu, "dmu"),
                                 _f(dlogvar, "dlogvar"), _f(partial, "partial"),
                                 _f(losses, "losses"), B, D, scale, _stream())
    if rc:
        _err(rc, "kl_fwd_bwd")
    add_launches(_KERNELS_PER_CALL["kl_fwd_bwd"])


def lc_fwd_bwd(z, bs, D, r0, r1, eta1, eta2, act_ws, partial, dz, losses):
    rc = load().sdvae_lc_fwd_bwd(_f(z, "z"), bs, D, r0, r1, eta1, eta2,
                                 _chk(act_ws, torch.uint8, "act_ws"), _f(partial, "partial"),
                                 _f(dz, "dz"), _f(losses, "losses"), _stream())
    if rc:
        _err(rc, "lc_fwd_bwd")
    add_launches(_KERNELS_PER_CALL["lc_fwd_bwd"])


def total_loss(losses, w_kl, w_lc, w_lap, w_cls=0.0):
    rc = load().sdvae_total_loss(_f(losses, "losses"), w_kl, w_lc, w_lap, w_cls, _stream())
    if rc:
        _err(rc, "total_loss")
    add_launches(_KERNELS_PER_CALL["total_loss"])


def adam_tick(step_dev):
    rc = load().sdvae_adam_tick(_i(step_dev, "step"), _stream())
    if rc:
        _err(rc, "adam_tick")
    add_launches(_KERNELS_PER_CALL["adam_tick"])


def adam_step(p, grad, m, v, step_dev, step_host, lr, b1, b2, eps, wd, gscale=1.0):
    rc = load().sdvae_adam_step(_f(p, "p"), _f(grad, "grad"), _f(m, "m"), _f(v, "v"), p.numel(),
                                None if step_dev is None else _i(step_dev, "step"), step_host,
                                lr, b1, b2, eps, wd, gscale, _stream())
    if rc:
        _err(rc, "adam_step")
    add_launches(_KERNELS_PER_CALL["adam_step"])
