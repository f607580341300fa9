"""ctypes binding of libp2c.so (the C-ABI declared in include/point2cyl.h).

There is no fallback: if the shared library is missing or a call returns non-zero this raises.
The product path never substitutes torch ops or the CPU oracle for a kernel.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libp2c.so")

c_f32p = C.c_void_p
c_i64p = C.c_void_p
c_i32p = C.c_void_p
c_f64p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f32 = C.c_float
vp = C.c_void_p

class BnFold(C.Structure):
    """include/point2cyl.h: p2c_bn_fold (a BatchNorm whose finalisation is deferred to the consuming kernel)."""
    _fields_ = [("stats", C.c_void_p), ("count", C.c_int64), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("eps", C.c_float), ("momentum", C.c_float), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("scale_out", C.c_void_p), ("shift_out", C.c_void_p),
                ("mean_out", C.c_void_p), ("invstd_out", C.c_void_p), ("C", C.c_int)]


bnp = C.POINTER(BnFold)

# name -> argtypes; restype is int unless noted.  Kept in the same order as include/point2cyl.h.
SIGNATURES = {
    "p2c_version": [],
    "p2c_arch": [],
    "p2c_set_sm_budget": [i32],
    "p2c_set_pdl": [i32],
    "p2c_fps": [c_f32p, c_i64p, i32, i32, i32, c_i64p, c_f32p, vp],
    "p2c_ball_query": [c_f32p, c_f32p, i32, i32, i32, f32, i32, c_i64p, vp],
    "p2c_group": [c_f32p, c_f32p, i64, c_f32p, c_i64p, i32, i32, i32, i32, i32, c_f32p, i64, vp],
    "p2c_sa_first_layer": [c_f32p, c_f32p, c_i64p, c_f32p, i64, c_f32p, i64, c_f32p, i32, i32, i32, i32, i32, c_f32p,
                           i64, c_f64p, vp],
    "p2c_linear_small": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, i64, i32, i32, i32, vp],
    "p2c_linear_group_bias": [c_f32p, i64, c_f32p, i64, c_f32p, i32, c_f32p, c_f32p, bnp, c_f32p, i64, i32, i32, i32,
                              c_f64p, vp],
    "p2c_group_moments_size": [],
    "p2c_group_moments": [c_f32p, c_f32p, c_i64p, i32, i32, i32, i32, c_f64p, vp],
    "p2c_sa_xyz_stats": [c_f64p, i64, c_f32p, i64, c_f32p, i32, c_f64p, vp],
    "p2c_sa_xyz_linear": [c_f32p, c_f32p, c_i64p, i32, i32, i32, i32, c_f32p, i64, c_f32p, i32, c_f32p, c_f32p, bnp,
                          c_f64p, c_f32p, c_f32p, i32, c_f32p, i64, c_f64p, i32, c_f32p, c_f32p, vp],
    "p2c_sa_stack_fused": [c_f32p, c_f32p, c_i64p, i32, i32, i32, i32, c_f32p, i64, c_f32p, bnp, c_f32p, c_f32p, bnp,
                           c_f32p, c_f32p, bnp, i32, i32, i32, c_f32p, i64, vp],
    "p2c_linear": [c_f32p, i64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, i64, c_f32p, i64, i32, i32, i32,
                   c_f64p, i32, c_f32p, c_f32p, i32, c_f32p, i64, bnp, vp],
    "p2c_split_tf32": [c_f32p, i32, i32, c_f32p, i64, vp],
    "p2c_split_tf32_multi": [vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "p2c_linear_act": [c_f32p, i64, c_f32p, i64, c_f32p, i32, i32, i32, i32, f32, f32, c_f32p, i64, c_f32p, i64, c_f32p,
                       i64, vp],
    "p2c_linear_act_bwd": [c_f32p, i64, c_f32p, i64, i32, i32, i32, i32, f32, f32, c_f32p, i64, c_f32p, i64, c_f32p, i64,
                           c_f32p, i64, vp],
    "p2c_cast_bf16": [c_f32p, i32, i32, vp, i64, vp],
    "p2c_linear_path": [i64, i32, i32, i32, i32, i32, i32, i32],
    "p2c_debug_set_timeline": [vp],
    "p2c_head_masked": [c_f32p, i64, c_f32p, c_f32p, c_f32p, c_i64p, c_f32p, c_f32p, c_f32p, i64, i32, i32, i32, i32,
                        bnp, i32, vp],
    "p2c_bn_finalize": [c_f64p, i64, c_f32p, c_f32p, f32, f32, i32, c_f32p, c_f32p, c_f32p, c_f32p,
                        c_f32p, c_f32p, i32, vp],
    "p2c_bn_relu_apply": [c_f32p, i64, c_f32p, c_f32p, c_f32p, i64, i64, i32, bnp, vp],
    "p2c_pool_bn_relu": [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, i64, i64, i32, bnp, vp],
    "p2c_three_nn_interp": [c_f32p, c_f32p, c_f32p, i64, i32, i32, i32, i32, c_f32p, i64, c_i64p, c_f32p, vp],
    "p2c_three_nn_search": [c_f32p, c_f32p, i32, i32, i32, c_i64p, c_f32p, vp],
    "p2c_three_nn_gather": [c_f32p, i64, c_i64p, c_f32p, i32, i32, i32, i32, c_f32p, i64, vp],
    "p2c_segfit_stats_stride": [i32],
    "p2c_segfit_stats": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, c_i64p, c_i64p, i32, i32, i32, c_f32p,
                         i64, c_f32p, vp],
    "p2c_segfit_stats_w": [c_f32p, i64, i32, c_f32p, i64, i64, c_f32p, i64, i64, c_f32p, c_f32p, c_i64p, c_i64p, i32, i32,
                           i32, c_f32p, i64, c_f32p, vp],
    "p2c_segfit_cost": [c_f32p, i32, i32, c_f32p, c_i32p, vp],
    "p2c_hungarian": [c_f32p, c_i32p, i32, i32, c_i64p, vp],
    "p2c_bb_loss": [c_f32p, i64, c_i64p, c_i64p, c_i32p, i32, i32, i32, c_f32p, i64, c_f32p, vp],
    "p2c_loss_finalize": [c_f32p, c_f32p, c_i64p, c_i32p, c_f32p, c_f32p, i32, i32, i32, i32,
                          C.POINTER(C.c_float), c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, vp],
    "p2c_eig3x3_smallest": [c_f32p, i32, c_f32p, c_f32p, vp],
    "p2c_square_distance": [c_f32p, c_f32p, i32, i32, i32, c_f32p, vp],
    "p2c_gather_rows": [c_f32p, i64, c_i64p, i32, i32, i32, i32, c_f32p, vp],
    "p2c_segment_lists": [c_i64p, c_i64p, i32, i32, i32, i32, c_i32p, c_i32p, vp],
    "p2c_sketch_project": [c_f32p, c_f32p, i32, i32, i32, i32, c_i32p, c_i32p, c_i64p, c_f32p, c_f32p, f32, c_f32p,
                           c_f32p, c_f32p, c_f32p, c_i32p, c_f32p, vp],
    "p2c_sketch_project_bwd": [c_f32p, c_i32p, c_f32p, i32, i32, i32, i32, c_f32p, vp],
    "p2c_extrusion_extents": [c_f32p, i32, i32, i32, i32, c_i32p, c_i32p, c_i64p, c_f32p, c_f32p, c_f32p, c_f32p, vp],
    "p2c_hard_w_encoding": [c_f32p, i64, i64, i32, i32, i32, c_f32p, f32, c_f32p, c_i64p, vp],
    "p2c_normal_angle": [c_f32p, c_f32p, i32, i32, f32, c_f32p, c_f32p, vp],
    "p2c_loss_backward_coef": [c_f32p, c_i64p, c_i32p, c_f32p, c_f32p, c_f32p, i32, i32, i32, i32, c_f32p, vp],
    "p2c_segfit_backward": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, c_i64p, c_i64p, c_f32p, c_i64p, c_i32p, c_f32p,
                            i32, i32, i32, c_f32p, i64, c_f32p, i64, vp],
    "p2c_segfit_backward_w": [c_f32p, i64, i32, c_f32p, i64, i64, c_f32p, i64, i64, c_f32p, c_f32p, c_i64p, c_f32p,
                              i32, i32, i32, c_f32p, i64, c_f32p, c_f32p, vp],
    "p2c_eig3x3_backward": [c_f32p, c_f32p, i32, c_f32p, vp],
    "p2c_bn_bwd_reduce": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, i64, i32, c_f64p, vp],
    "p2c_pool_bwd_reduce": [c_f32p, i64, c_f32p, c_f32p, c_f32p, c_f32p, i64, i32, c_f64p, vp],
    "p2c_bn_bwd_coef": [c_f64p, i64, c_f32p, c_f32p, c_f32p, i32, c_f32p, c_f32p, c_f32p, i32, vp],
    "p2c_bn_bwd_apply": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, c_f32p, i64, i32, c_f32p, i64, vp],
    "p2c_pool_bwd_apply": [c_f32p, i64, c_f32p, c_f32p, c_f32p, i64, c_f32p, c_f32p, c_f32p, i64, i32, i32, c_f32p,
                           i64, vp],
    "p2c_bn_bwd_apply_fused": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, c_f64p, i64, c_f32p, c_f32p, c_f32p, i32,
                               c_f32p, c_f32p, i64, i32, c_f32p, i64, vp],
    "p2c_pool_bwd_apply_fused": [c_f32p, i64, c_f32p, c_f32p, c_f32p, i64, c_f32p, c_f32p, c_f64p, i64, c_f32p, c_f32p,
                                 c_f32p, i32, c_f32p, c_f32p, i64, i32, i32, c_f32p, i64, vp],
    "p2c_wgrad": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, c_f32p, i32, i64, i32, i32, c_f32p, i64, c_f32p, i32, vp],
    "p2c_wgrad_path": [i64, i64, i64, i32, i32, i32],
    "p2c_sa_first_bwd": [c_f32p, i64, c_f32p, c_f32p, c_i64p, i32, i32, i32, i32, i32, c_f32p, i64, c_f32p, i64,
                         c_f32p, vp],
    "p2c_group_bwd": [c_f32p, i64, c_i64p, i32, i32, i32, i32, i32, c_f32p, i64, vp],
    "p2c_three_nn_interp_bwd": [c_f32p, i64, c_i64p, c_f32p, i32, i32, i32, i32, c_f32p, i64, vp],
    "p2c_head_bwd": [c_f32p, i64, c_f32p, c_i64p, c_f32p, i32, i32, i32, i32, c_f32p, i64, c_f32p, i64, c_f32p, c_f32p,
                     c_f32p, i64, vp],
    "p2c_igr_add_latent": [c_f32p, c_f32p, i64, i32, i32, c_f32p, i64, c_f32p, i64, i32, f32, vp],
    "p2c_igr_scale_cols": [c_f32p, i64, c_f32p, i64, i32, c_f32p, i64, vp],
    "p2c_igr_rowdots": [c_f32p, i64, i64, i32, c_f32p, i64, i64, i32, c_f32p, f32, c_f32p, i64, i32, vp],
    "p2c_igr_loss_terms": [c_f32p, c_f32p, c_f32p, c_f32p, i32, i32, i32, c_f32p, vp],
    "p2c_igr_colsums": [c_f32p, i64, i64, i32, c_f32p, i64, i32, f32, c_f32p, i64, vp],
    "p2c_igr_seed_delta": [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, i64, i32, c_f32p, i64, vp],
    "p2c_igr_latent_grad": [c_f32p, i64, i32, i32, i32, f32, c_f32p, c_f32p, i32, vp],
    "p2c_igr_loss_terms_bwd": [c_f32p, c_f32p, c_f32p, c_f32p, i32, i32, i32, c_f32p, c_f32p, c_f32p, c_f32p, vp],
    "p2c_adam_step": [c_f32p, c_f32p, c_f32p, c_f32p, i64, f32, f32, f32, f32, f32, i32, f32, vp],
}

P2C_ERRORS = {-1: "P2C_EINVAL (bad size / null pointer)", -2: "P2C_EUNSUPPORTED (shape not covered)",
              -3: "P2C_EALIGN (alignment)"}

PREC_FP32, PREC_3XTF32, PREC_BF16 = 0, 1, 2

_lib = None


class P2CError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libp2c.so (once).  Raises if it has not been built: there is no CPU/torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise P2CError(
            f"{LIB_PATH} not found: build it with `python -m point2cyl_b200.build` "
            "(or __graft_entry__.build()). point2cyl_b200 has no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "p2c_arch" else C.c_int
    if os.environ.get("P2C_PDL") is not None:          # A/B switch for measurements (default: on)
        lib.p2c_set_pdl(int(os.environ["P2C_PDL"] != "0"))
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise P2CError(f"{what}: {P2C_ERRORS.get(rc, rc)}")
    raise P2CError(f"{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")


_alive: list = []


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor is kept alive until the next `call()` has returned, so a
    temporary made inline (`ptr(x.contiguous())`) cannot be released between taking its address and the launch that
    reads it; after the launch the allocator's stream ordering protects it."""
    if t is None:
        return None
    _alive.append(t)
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def need_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise P2CError("point2cyl_b200 kernels need CUDA tensors; there is no CPU path "
                           "(the CPU oracle under oracle/ is test infrastructure only)")


# kernels launched by one call of each entry point (the `gpu_launches` claim of bench.py)
LAUNCHES_PER_CALL = {
    "p2c_split_tf32": 1, "p2c_split_tf32_multi": 1, "p2c_cast_bf16": 1, "p2c_sa_first_layer": 1, "p2c_head_masked": 1, "p2c_fps": 1, "p2c_ball_query": 1, "p2c_group": 1, "p2c_linear": 1, "p2c_bn_finalize": 1,
    "p2c_bn_relu_apply": 1, "p2c_pool_bn_relu": 1, "p2c_three_nn_interp": 1, "p2c_three_nn_search": 1, "p2c_three_nn_gather": 1, "p2c_segfit_stats": 2, "p2c_segfit_stats_w": 2,
    "p2c_segfit_cost": 1, "p2c_hungarian": 1, "p2c_bb_loss": 2, "p2c_loss_finalize": 2, "p2c_eig3x3_smallest": 1,
    "p2c_square_distance": 1, "p2c_gather_rows": 1, "p2c_segment_lists": 1, "p2c_sketch_project": 1, "p2c_sketch_project_bwd": 1,
    "p2c_extrusion_extents": 1, "p2c_hard_w_encoding": 1, "p2c_normal_angle": 1,
    "p2c_bn_bwd_reduce": 1, "p2c_pool_bwd_reduce": 1, "p2c_bn_bwd_coef": 1, "p2c_bn_bwd_apply": 1,
    "p2c_pool_bwd_apply": 1, "p2c_wgrad": 1, "p2c_sa_first_bwd": 1, "p2c_group_bwd": 1, "p2c_three_nn_interp_bwd": 1, "p2c_head_bwd": 1,
    "p2c_linear_small": 1, "p2c_linear_group_bias": 1, "p2c_group_moments": 1, "p2c_sa_xyz_stats": 1, "p2c_sa_xyz_linear": 1,
    "p2c_adam_step": 1, "p2c_loss_backward_coef": 1, "p2c_segfit_backward": 1, "p2c_segfit_backward_w": 1,
    "p2c_eig3x3_backward": 1,
}
launch_count = 0
_profile = None  # None, or list of (name, tag, start_event, end_event)
_profile_tag = ""


def call(name: str, *args) -> None:
    """Invoke one C-ABI entry point, count its kernel launches, optionally time it with CUDA events
    on the launching stream."""
    global launch_count
    fn = getattr(load(), name)
    if _profile is None:
        rc = fn(*args)
    else:
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        _profile.append((name, _profile_tag, s, e))
    _alive.clear()
    check(rc, name)
    launch_count += LAUNCHES_PER_CALL.get(name, 1)


def profile_start() -> None:
    global _profile
    _profile = []


def profile_stop():
    """-> list of (entry point, tag, milliseconds); synchronises."""
    global _profile
    torch.cuda.synchronize()
    out = [(n, t, s.elapsed_time(e)) for n, t, s, e in (_profile or [])]
    _profile = None
    return out


def set_tag(tag: str) -> None:
    global _profile_tag
    _profile_tag = tag
