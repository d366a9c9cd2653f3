"""ctypes binding of libvanerf_b200.so (include/vanerf_b200.h).

The product path loads the nvcc-built CUDA library only and fails loudly when it is missing: there is no CPU
fallback.  (`Lib(path, emulated=True)` exists for tests/ only, which point it at the host-emulation build of the
same kernel sources to check kernel logic in a container without a GPU.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VANERF_B200_LIB: developer override (e.g. the cycle-trace build tools/build.py trace makes); same ABI, same kernels
LIB_PATH = os.environ.get("VANERF_B200_LIB") or os.path.join(_HERE, "libvanerf_b200.so")

MAX_VIEWS = 4            # VANERF_MAX_VIEWS (fp32 path)
MAX_VIEWS_BF16 = 3       # VANERF_MAX_VIEWS_BF16 (tensor-core path)
RAY_STRIDE = 8
FP32, BF16 = 0, 1


class VLinear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("out_dim", C.c_int32), ("in_dim", C.c_int32)]


class VWeights(C.Structure):
    _fields_ = [("geo_at", VLinear * 2), ("geo_f", VLinear * 2), ("geo8_at", VLinear * 2), ("geo8_f", VLinear * 2),
                ("mlp", VLinear * 4), ("post", VLinear * 3), ("compress", VLinear),
                ("tex_at", VLinear * 2), ("tex_f", VLinear * 2),
                ("ray", VLinear * 2), ("base", VLinear * 2), ("vis1", VLinear * 2), ("vis2", VLinear * 2),
                ("outl", VLinear * 3), ("ani_al", C.c_float), ("sigmoid_beta", C.c_float)]


class VFrame(C.Structure):
    _fields_ = [("n_views", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("znear", C.c_float), ("zfar", C.c_float), ("z_range", C.c_float),
                ("KRT", C.c_void_p), ("extrin", C.c_void_p), ("src_cam_pos", C.c_void_p), ("kpt3d", C.c_void_p),
                ("verts", C.c_void_p), ("faces", C.c_void_p), ("n_verts", C.c_int32), ("n_faces", C.c_int32),
                ("img", C.c_void_p), ("fg_mask", C.c_void_p),
                ("feat_geo0", C.c_void_p), ("g0_h", C.c_int32), ("g0_w", C.c_int32),
                ("feat_geo1", C.c_void_p), ("g1_h", C.c_int32), ("g1_w", C.c_int32),
                ("feat_tex", C.c_void_p), ("t_h", C.c_int32), ("t_w", C.c_int32),
                ("vert_gfeat", C.c_void_p)]


class VConvStack(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv0", "ln1_w", "ln1_b", "conv3", "ln4_w", "ln4_b")]


class VGfeatWeights(C.Structure):
    _fields_ = [("img", VConvStack), ("tex", VConvStack), ("gt", VConvStack)]


class VTarget(C.Structure):
    _fields_ = [("inv_K", C.c_float * 9), ("R", C.c_float * 9), ("cam_pos", C.c_float * 3),
                ("znear", C.c_float), ("zfar", C.c_float), ("bounds", C.c_float * 6)]


_P, _I, _I64 = C.c_void_p, C.c_int32, C.c_int64

# name -> (restype, argtypes); every symbol declared in include/vanerf_b200.h
PROTOTYPES = {
    "vanerf_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "vanerf_ctx_destroy": (None, [_P]),
    "vanerf_status_str": (C.c_char_p, [C.c_int]),
    "vanerf_last_error": (C.c_char_p, [_P]),
    "vanerf_sm_count": (C.c_int, [_P]),
    "vanerf_load_weights": (C.c_int, [_P, C.POINTER(VWeights), _P]),
    "vanerf_frame_setup": (C.c_int, [_P, C.POINTER(VFrame), _P, _P]),
    "vanerf_global_vertex_feature": (C.c_int, [_P, C.POINTER(VGfeatWeights), _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "vanerf_sample_rays": (C.c_int, [_P, C.POINTER(VTarget), _P, _I, _P, _I, _P, _P, _P]),
    "vanerf_sample_rays_t": (C.c_int, [_P, C.POINTER(VTarget), _P, _I, _P, _I, _I, _P, _P, _P]),
    "vanerf_geom_query": (C.c_int, [_P, C.POINTER(VTarget), _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "vanerf_shade": (C.c_int, [_P, C.c_int, C.POINTER(VTarget), _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "vanerf_shade_debug": (C.c_int, [_P, C.POINTER(VTarget), _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vanerf_composite": (C.c_int, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "vanerf_importance": (C.c_int, [_P, _P, _P, _I, _I, _P, _I, _I, _P, _P, _P]),
    "vanerf_importance_mid": (C.c_int, [_P, _P, _P, _I, _I, _P, _I, _I, _P, _P]),
    "vanerf_set_reuse_coarse": (C.c_int, [_P, C.c_int]),
    "vanerf_set_reuse_geometry": (C.c_int, [_P, C.c_int]),
    "vanerf_render_rays": (C.c_int, [_P, C.c_int, C.POINTER(VTarget), _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "vanerf_query_points": (C.c_int, [_P, C.c_int, C.POINTER(VTarget), _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    "vanerf_feat_sample": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P]),
    "vanerf_knn1": (C.c_int, [_P, _P, _I, _P, _I, _P, _P]),
    "vanerf_dense": (C.c_int, [_P, _P, _I, _I, _P, _P, _I, _I, _P, _P]),
    "vanerf_rel_z_decay": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, C.c_float, C.c_float, _P, _P]),
    "vanerf_project_samples": (C.c_int, [_P, C.POINTER(VTarget), _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "vanerf_feat_sample_bwd": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P]),
    "vanerf_composite_beta": (C.c_int, [_P, _P, _P, _P, _I, _I, C.c_float, _P, _P, _P, _P, _P, _P]),
    "vanerf_composite_bwd": (C.c_int, [_P, _P, _P, _P, _I, _I, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "vanerf_timing_enable": (C.c_int, [_P, C.c_int]),
    "vanerf_timing_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(_I64), C.c_int]),
    "vanerf_scratch_bytes": (C.c_size_t, [_P, _I, _I]),
    "vanerf_launch_count": (_I64, [_P]),
    "vanerf_shade_debug_bf16": (C.c_int, [_P, C.POINTER(VTarget), _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vanerf_tc_error": (C.c_int, [_P]),
    "vanerf_tc_check": (C.c_int, [_P, _P]),
    "vanerf_tc_profile": (C.c_int, [_P, _P, _I]),
    "vanerf_tc_selftest": (C.c_int, [_P, _P, _P, _I, _I, _P, _P]),
    "vanerf_tc_mma_probe": (C.c_int, [_P, _I, _I, _I, _I, _I, _P]),
    "vanerf_tc_program_check": (C.c_int, []),
}


class VanerfError(RuntimeError):
    pass


class Lib:
    def __init__(self, path: str = LIB_PATH, emulated: bool = False):
        if not os.path.exists(path):
            raise VanerfError(
                f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "vanerf_b200 has no CPU fallback.")
        self.path, self.emulated = path, emulated
        self.dll = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(self.dll, name)      # AttributeError = symbol missing from the build
            fn.restype, fn.argtypes = res, args

    def check(self, ctx, status: int, what: str):
        if status != 0:
            msg = self.dll.vanerf_status_str(status).decode()
            detail = self.dll.vanerf_last_error(ctx).decode() if ctx else ""
            raise VanerfError(f"{what}: {msg} ({status}) {detail}")


_lib = None


def get_lib() -> Lib:
    """The CUDA library (product path)."""
    global _lib
    if _lib is None:
        _lib = Lib(LIB_PATH, emulated=False)
    return _lib
