"""ctypes binding of the C-ABI library ``libtricolo_b200.so`` (include/tricolo_b200.h).

The library is the only compute path of this package: if it cannot be loaded
the import fails loudly — there is no PyTorch / CPU fallback (BASELINE.json
north_star: "no Triton, no multi-backend dispatch and no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libtricolo_b200.so"

TCL_OP_F16, TCL_OP_BF16 = 0, 1
TCL_DT_F32, TCL_DT_F16, TCL_DT_BF16, TCL_DT_F64 = 0, 1, 2, 3
KERNEL_IDS = {
    "l2norm_fwd": 0, "cast16": 1, "transpose16": 2, "ntxent_fwd": 3, "fwd_reduce": 4, "fwd_finalize": 5,
    "ntxent_bwd": 6, "l2norm_bwd": 7, "sim_gemm": 8, "topk_rank": 9, "gather_gt": 10, "topk_merge": 11, "sim_topk_fused": 12, "gather_sum": 13, "peer_sum": 14, "ntxent_g": 15, "rank_metrics": 16,
    "ntxent_small_fwd": 17, "ntxent_small_bwd": 18, "ntxent_raw_fwd": 19, "ntxent_raw_bwd": 20,
}

_DTYPE_CODE = {
    torch.float32: TCL_DT_F32,
    torch.float16: TCL_DT_F16,
    torch.bfloat16: TCL_DT_BF16,
    torch.float64: TCL_DT_F64,
}
_OP_TORCH = {TCL_OP_F16: torch.float16, TCL_OP_BF16: torch.bfloat16}


class TricoloB200Error(RuntimeError):
    """Non-zero return code from the C ABI."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"tricolo_b200 error {code}: {msg}")
        self.code = code


class BwdSegment(C.Structure):
    _fields_ = [
        ("z_other", C.c_void_p),
        ("z_other_t", C.c_void_p),
        ("lse2_self", C.c_void_p),
        ("lse2_other", C.c_void_p),
        ("grad_scale", C.c_void_p),
        ("w_self", C.c_float),
        ("w_other", C.c_float),
    ]


class BwdJob(C.Structure):
    _fields_ = [
        ("z_self", C.c_void_p),
        ("x_self", C.c_void_p),
        ("inv_norm", C.c_void_p),
        ("dx", C.c_void_p),
        ("n_segments", C.c_int32),
        ("reserved", C.c_int32),
        ("seg", BwdSegment * 2),
    ]


# symbol -> (restype, argtypes); the single source for the loader and for the
# CPU test that checks every declared symbol is exported.
_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "tcl_version": (_i, []),
    "tcl_last_error_string": (C.c_char_p, []),
    "tcl_l2norm_fwd": (_i, [_i, _pp, _i, _i64, _i64, _i64, _pp, _i64, _i, _pp, _f, _vp]),
    "tcl_l2norm_fwd_bcast": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, _pp, _i64, _i, _pp, _f, _vp]),
    "tcl_peer_sum_f32": (_i, [_i, _pp, _i64, _vp, _vp]),
    "tcl_shard_sync_bytes": (_sz, []),
    "tcl_shard_stats_bytes": (_sz, [_i, _i64, _i]),
    "tcl_l2norm_fwd_push": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, _i, _pp, _i64, _i, _pp, _f, _pp, _i, _vp]),
    "tcl_ntxent_fwd_sharded": (_i, [_i, _pp, _pp, _i64, _i64, _i64, _i64, _i, _i, _i, _f, _vp, _vp, _sz, _pp, _pp, _pp, _i,
                                    C.POINTER(C.c_int64), _vp]),
    "tcl_ntxent_finalize_sharded": (_i, [_i, _i64, _i64, _i, _i, _f, _f, _pp, _vp, _vp, _vp, _vp, _vp]),
    "tcl_copy_rows": (_i, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "tcl_cast_16bit": (_i, [_vp, _i, _i64, _i64, _i64, _vp, _i, _vp]),
    "tcl_triplet_workspace_bytes": (_sz, [_i64]),
    "tcl_triplet_fwd": (_i, [_vp, _vp, _i, _i64, _i64, _i64, _f, _vp, _vp, _vp, _sz, _vp]),
    "tcl_triplet_bwd": (_i, [_vp, _vp, _i, _i64, _i64, _i64, _f, _vp, _vp, _sz, _vp, _vp, _vp]),
    "tcl_transpose_16bit": (_i, [_i, _pp, _i64, _i64, _i64, _pp, _i64, _vp]),
    "tcl_gather_sum_cast16": (_i, [_i, _pp, _i, _i64, _i64, _i64, _vp, _i64, _vp, _i, _vp]),
    "tcl_ntxent_fwd_workspace_bytes": (_sz, [_i, _i64, _i64]),
    "tcl_ntxent_fwd": (_i, [_i, _pp, _pp, _i64, _i64, _i64, _i64, _i64, _i, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "tcl_ntxent_finalize": (_i, [_i, _i64, _i64, _i64, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tcl_ntxent_bwd_workspace_bytes": (_sz, [_i, _i64, _i64]),
    "tcl_ntxent_bwd": (_i, [_i, C.POINTER(BwdJob), _i64, _i64, _i64, _i64, _i64, _i64, _i, _i64, _i, _f, _f, _vp, _sz, _vp]),
    "tcl_ntxent_bwd_sharded_workspace_bytes": (_sz, [_i, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                                     _i64, _i64, _i64, _i]),
    "tcl_ntxent_bwd_sharded_recv_bytes": (_sz, [_i, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                                _i64, _i64, _i64, _i]),
    "tcl_ntxent_bwd_sharded_gemm": (_i, [_i, _pp, _i64, _i64, _i64, _i64, _i, _i, _i, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), _i, _f, _f, _vp, _vp, _vp, C.POINTER(C.c_uint8), _vp, _sz,
                                         _pp, _sz, _pp, _vp]),
    "tcl_ntxent_bwd_sharded_finish": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i64, _i, _i, _i, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), _vp, C.POINTER(C.c_uint8), _f, _vp, _vp, _vp, _pp, _vp]),
    "tcl_ntxent_loss_state_bytes": (_sz, [_i, _i, _i64, _i64]),
    "tcl_ntxent_loss_workspace_bytes": (_sz, [_i, _i, _i64, _i64]),
    "tcl_ntxent_loss_fwd": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 _i, _f, _f, _f, _vp, _sz, _vp, _sz, _vp, _vp]),
    "tcl_ntxent_loss_bwd": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 _i, _f, _f, _f, _vp, _vp, C.POINTER(C.c_uint8), _pp, _vp, _sz, _vp]),
    "tcl_ntxent_loss_fwd_total": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       _i, _f, _f, _f, _vp, _sz, _vp, _sz, _vp, _vp]),
    "tcl_ntxent_loss_bwd_total": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       _i, _f, _f, _f, _vp, _vp, _vp, C.POINTER(C.c_uint8), _pp, _vp, _sz, _vp]),
    "tcl_ntxent_raw_state_bytes": (_sz, [_i, _i64]),
    "tcl_ntxent_raw_workspace_bytes": (_sz, [_i, _i64]),
    "tcl_ntxent_raw_fwd": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                _f, _f, _vp, _sz, _vp, _sz, _vp, _vp]),
    "tcl_ntxent_raw_bwd": (_i, [_i, _pp, _i, _i64, _i64, _i64, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                _f, _f, _vp, _vp, _vp, C.POINTER(C.c_uint8), _pp, _vp]),
    "tcl_sim_gemm": (_i, [_vp, _vp, _i64, _i64, _i64, _i, _vp, _i64, _vp]),
    "tcl_topk_rank": (_i, [_vp, _i64, _i64, _i64, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tcl_gather_gt_sim": (_i, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    "tcl_topk_merge": (_i, [_vp, _vp, _i, _i64, _i, _vp, _vp, _vp]),
    "tcl_rank_metrics_workspace_bytes": (_sz, []),
    "tcl_rank_metrics": (_i, [_vp, _i64, _i, _vp, _vp, _sz, _vp]),
    "tcl_gt_sim_mma": (_i, [_vp, _vp, _i64, _i64, _i64, _i, _vp, _i64, _vp, _vp, _sz, _vp]),
    "tcl_sim_topk_fused_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "tcl_sim_topk_fused": (_i, [_vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "tcl_launch_count": (_i64, []),
    "tcl_debug_small_trace": (_i, [C.POINTER(C.c_uint64)]),
    "tcl_profile_enable": (_i, [_i]),
    "tcl_profile_read": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "tcl_debug_tmem_probe": (_i, [_vp, _vp]),
    "tcl_ntxent_bwd_needs_transpose": (_i, [_i64]),
    "tcl_debug_pc_trace": (_i, [C.POINTER(C.c_uint64), _i]),
    "tcl_debug_fwd_trace": (_i, [C.POINTER(C.c_uint64), _i]),
    "tcl_debug_gb_trace": (_i, [C.POINTER(C.c_uint64), _i]),
    "tcl_debug_max_clusters": (_i, [_i, C.POINTER(C.c_int)]),
}


def lib_path() -> Path:
    return Path(os.environ.get("TRICOLO_B200_LIB", str(_LIB_PATH)))


def _load() -> C.CDLL:
    path = lib_path()
    if not path.exists():
        raise ImportError(
            f"tricolo_b200: {path} not found. Build it with `make` (or "
            "`python -c 'import __graft_entry__ as g; g.build()'`) — there is no fallback path."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


LIB = _load()


def check(code: int) -> None:
    if code != 0:
        raise TricoloB200Error(code, LIB.tcl_last_error_string().decode("utf-8", "replace"))


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"tricolo_b200: unsupported dtype {t.dtype}") from None


def op_torch_dtype(op_format: int) -> torch.dtype:
    return _OP_TORCH[op_format]


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def ptr_array(tensors) -> "C.Array":
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def stream_ptr(device=None) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "tricolo_b200 runs on sm_100a only: got a tensor on "
                f"{t.device}; there is no CPU path in this package"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tricolo_b200: tensors on different devices ({dev} vs {t.device})")
    return dev


def launch_count() -> int:
    return int(LIB.tcl_launch_count())


def profile_enable(on: bool) -> None:
    check(LIB.tcl_profile_enable(1 if on else 0))


def profile_read() -> dict:
    """{kernel name: (total device ms, launches)} since the last profile_enable(True)."""
    out = {}
    for name, kid in KERNEL_IDS.items():
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(LIB.tcl_profile_read(kid, C.byref(ms), C.byref(n)))
        if n.value:
            out[name] = (ms.value, n.value)
    return out
