"""ctypes loader of libmetaasr_b200.so (the C ABI declared in include/metaasr_b200.h).

There is NO fallback: if the shared library is missing or no sm_100 GPU is visible, every
compute entry point raises.  The library is built in-tree by csrc/build.sh (see
__graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libmetaasr_b200.so"

_lib = None
_inited_devices = set()

c_p = C.c_void_p
c_i = C.c_int
c_i64 = C.c_int64
c_f = C.c_float
c_d = C.c_double
c_u64 = C.c_uint64
c_u32 = C.c_uint32
c_sz = C.c_size_t

# name -> argtypes, exactly the declarations of include/metaasr_b200.h (restype int unless noted)
SIGNATURES = {
    "masr_abi_version": [],
    "masr_init": [c_i],
    "masr_ctc_fwd_bwd": [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p, c_p, c_p,
                         c_p, c_sz, c_p],
    "masr_ctc_fwd_bwd_ex": [c_p, c_i, c_i, c_i, c_i64, c_i64, c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p, c_p, c_p,
                            c_p, c_sz, c_p],
    "masr_loss_mix": [c_p, c_p, c_f, c_p],
    "masr_cast_pad2d": [c_p, c_i, c_i64, c_p, c_i, c_i64, c_i, c_i, c_p],
    "masr_ctc_debug_enable": [c_i],
    "masr_ctc_debug_read": [c_p],
    "masr_gemm": [c_p, c_i, c_i64, c_i64, c_p, c_i, c_i64, c_i64, c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_umma_gemm": [c_p, c_i64, c_i, c_p, c_i64, c_i, c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_umma_gemm_ex": [c_p, c_i64, c_i, c_p, c_i64, c_i, c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "masr_umma_gemm_pair": [c_p, c_i64, c_i, c_p, c_i64, c_i, c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "masr_gemm_set_pair_mode": [c_i],
    "masr_gemm_group_begin": [],
    "masr_gemm_group_end": [c_p],
    "masr_gemm_group_last": [c_p, c_p],
    "masr_attn_set_small_lq": [c_i],
    "masr_gemm_set_stage_cap": [c_i],
    "masr_umma_gemm_tn": [c_p, c_i64, c_p, c_i64, c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i, c_p],
    "masr_umma_conv3x3_fwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_umma_conv3x3_dgrad": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_umma_conv1_fwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "masr_umma_conv1_wgrad": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "masr_umma_conv3x3_wgrad": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_conv1_fwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_conv1_wgrad": [c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "masr_im2col3x3": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_col2im3x3": [c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_p],
    "masr_conv_w_prep": [c_p, c_p, c_i, c_i, c_i, c_p],
    "masr_conv_w_prep_t": [c_p, c_p, c_i, c_i, c_i, c_p],
    "masr_conv_w_unprep_add": [c_p, c_p, c_i, c_i, c_p],
    "masr_maxpool2x2_fwd": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_maxpool2x2_bwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_maxpool2x2_fwd_code": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_maxpool2x2_bwd_code": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_relu_bwd": [c_p, c_p, c_i, c_i64, c_f, c_p],
    "masr_attn_fwd": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i,
                      c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_f, c_u64, c_u32, c_p],
    "masr_attn_bwd": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_p,
                      c_p, c_i64, c_p, c_i64, c_p, c_i64, c_i,
                      c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_f, c_u64, c_u32, c_p],
    "masr_umma_attn_fwd": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p,
                           c_i, c_i, c_i, c_i, c_p, c_i, c_f, c_u64, c_u32, c_p],
    "masr_umma_attn_fwd_cached": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p,
                                  c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_f, c_u64, c_u32, c_p],
    "masr_attn_fwd_cached": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i,
                             c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_f, c_u64, c_u32, c_p],
    "masr_umma_attn_bwd": [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_p, c_p,
                           c_p, c_i64, c_p, c_i64, c_p, c_i64,
                           c_i, c_i, c_i, c_i, c_p, c_i, c_f, c_u64, c_u32, c_i, c_p, c_p],
    "masr_add_layernorm_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_u64, c_u32, c_p],
    "masr_add_layernorm_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_u64, c_u32, c_p],
    "masr_add_pe_dropout": [c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_u64, c_u32, c_p],
    "masr_embed_pe_fwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_u64, c_u32, c_p],
    "masr_embed_bwd": [c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_f, c_u64, c_u32, c_p],
    "masr_dropout": [c_p, c_i, c_i64, c_f, c_u64, c_u32, c_p],
    "masr_colsum_add": [c_p, c_i, c_i64, c_p, c_i, c_i, c_p],
    "masr_cast": [c_p, c_i, c_p, c_i, c_i64, c_p],
    "masr_permute_cf": [c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "masr_ls_ce_fwd_bwd": [c_p, c_p, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p, c_i, c_i64, c_p],
    "masr_set_seed_ptr": [c_p],
    "masr_seed_bump": [c_p, c_u64, c_p],
    "masr_mt_sumsq": [c_p, c_i64, c_p, c_i, c_p],
    "masr_mt_clip_sgd": [c_p, c_p, c_p, c_i64, c_p, c_f, c_f, c_f, c_i, c_i, c_p],
    "masr_mt_clip_sgd_ex": [c_p, c_p, c_p, c_i64, c_p, c_f, c_f, c_f, c_i, c_i, c_p, c_i, c_p],
    "masr_mt_copy_cast": [c_p, c_p, c_p, c_i64, c_p],
    "masr_prep_weights": [c_p, c_p, c_i64, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "masr_mt_clip": [c_p, c_i64, c_p, c_f, c_p],
    "masr_mt_accumulate": [c_p, c_p, c_i64, c_p, c_f, c_p],
    "masr_mt_reptile_delta": [c_p, c_p, c_p, c_i64, c_p],
    "masr_mt_adam": [c_p, c_p, c_p, c_p, c_i64, c_f, c_f, c_f, c_f, c_f, c_d, c_d, c_p, c_p, c_f, c_p],
    "masr_mt_axpy": [c_p, c_p, c_f, c_i64, c_p],
    "masr_nvls_allreduce_f32": [c_p, c_i64, c_i64, c_p],
    "masr_nvls_reduce_adam": [c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_f, c_f, c_f, c_f, c_f, c_d, c_d, c_p],
}
SPECIAL_RESTYPE = {"masr_last_error": C.c_char_p, "masr_ctc_workspace_bytes": c_sz}
EXTRA = {"masr_last_error": [], "masr_ctc_workspace_bytes": [c_i, c_i, c_i, c_i]}


class GemmEpilogue(C.Structure):
    """masr_gemm_epilogue of include/metaasr_b200.h."""
    _fields_ = [("rowsum", c_p), ("mask", c_p), ("ldmask", c_i64), ("mask_scale", c_f), ("p_drop", c_f),
                ("seed", c_u64), ("site", c_u32), ("dot_src", c_p), ("lddot", c_i64), ("dot_out", c_p),
                ("dot_L", c_i), ("dot_H", c_i)]


class ConvPrepJob(C.Structure):
    """masr_conv_prep_job of include/metaasr_b200.h."""
    _fields_ = [("w", c_p), ("wp", c_p), ("wpt", c_p), ("Cout", c_i), ("Cin", c_i)]


class MetaASRLibraryError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every declared symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("METAASR_B200_LIB", LIB_PATH))
    if not path.exists():
        raise MetaASRLibraryError(
            f"{path} not found: build it with metaasr_crossaccent_b200/csrc/build.sh "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, argtypes in {**SIGNATURES, **EXTRA}.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = SPECIAL_RESTYPE.get(name, c_i)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().masr_last_error()
        raise MetaASRLibraryError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def init(device_index: int):
    """masr_init: verifies an sm_100 device.  Must succeed before any compute call."""
    lib = load()
    if device_index not in _inited_devices:
        check(lib.masr_init(int(device_index)), "masr_init")
        _inited_devices.add(device_index)
    return lib
