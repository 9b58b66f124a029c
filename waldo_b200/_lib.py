"""ctypes binding of libwaldo_b200.so -- the C ABI declared in include/waldo_b200.h.

This is the reference-side binding a maintainer of 16lemoing/waldo would add: the
structures below mirror the header field by field (tests/test_abi.py parses the header and checks).
There is NO fallback: if the sm_100a library cannot be built/loaded, or a tensor is not a contiguous CUDA
tensor of the expected dtype, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# WALDO_B200_LIB points at an alternative build of the SAME sources (kernel experiments); it must still be an sm_100a
# device library -- load() checks waldo_has_device_code() either way.
LIB_PATH = os.environ.get("WALDO_B200_LIB") or os.path.join(_HERE, "libwaldo_b200.so")

c_float_p = C.c_void_p
c_void_p = C.c_void_p


class TpsFwd(C.Structure):
    _fields_ = [("n", C.c_int), ("N", C.c_int), ("P", C.c_int),
                ("inverse_kernel", c_void_p), ("tgt_grid_repr", c_void_p), ("pts", c_void_p),
                ("mapping", c_void_p), ("grid", c_void_p)]


class TpsBwd(C.Structure):
    _fields_ = [("n", C.c_int), ("N", C.c_int), ("P", C.c_int),
                ("inverse_kernel", c_void_p), ("tgt_grid_repr", c_void_p), ("dgrid", c_void_p),
                ("chunks", C.c_int), ("partial", c_void_p), ("dpts", c_void_p)]


class InvWarpFwd(C.Structure):
    _fields_ = [("n", C.c_int), ("Hs", C.c_int), ("Ws", C.c_int), ("Ht", C.c_int), ("Wt", C.c_int),
                ("niter", C.c_int), ("erode", C.c_int),
                ("fwd_grid", c_void_p), ("id_src", c_void_p), ("id_tgt", c_void_p), ("gauss", c_void_p),
                ("out", c_void_p), ("field", c_void_p), ("winner", c_void_p), ("level", c_void_p),
                ("eroded", c_void_p), ("val", c_void_p), ("bbox", c_void_p)]


class InvWarpBwd(C.Structure):
    _fields_ = [("n", C.c_int), ("Hs", C.c_int), ("Ws", C.c_int), ("Ht", C.c_int), ("Wt", C.c_int), ("niter", C.c_int),
                ("gauss", c_void_p), ("dout", c_void_p), ("field", c_void_p), ("winner", c_void_p),
                ("level", c_void_p), ("eroded", c_void_p), ("bbox", c_void_p), ("gval", c_void_p), ("inv_sw", c_void_p),
                ("gdisp", c_void_p), ("dfwd_grid", c_void_p)]


class Geom(C.Structure):
    _fields_ = [("B", C.c_int), ("T", C.c_int), ("Tw", C.c_int), ("Tc", C.c_int), ("Tp", C.c_int),
                ("No", C.c_int), ("Nl", C.c_int), ("C", C.c_int),
                ("H", C.c_int), ("W", C.c_int), ("Hd", C.c_int), ("Wd", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int),
                ("flags", C.c_int), ("min_cls", C.c_float)]


class DecodeFwd(C.Structure):
    _fields_ = [("g", Geom),
                ("input", c_void_p), ("tgt_grid_obj", c_void_p), ("src_grid_obj", c_void_p),
                ("tgt_grid_bg", c_void_p), ("src_grid_bg", c_void_p), ("occ", c_void_p),
                ("obj_alpha", c_void_p), ("bg_alpha", c_void_p), ("cls", c_void_p),
                ("ctx_ts", c_void_p), ("pred_ts", c_void_p), ("xs_hd", c_void_p), ("ys_hd", c_void_p),
                ("a_lo", c_void_p), ("prof_part", c_void_p), ("prof_ctas", C.c_int), ("prof_sum", c_void_p),
                ("prof_p", c_void_p), ("lyt_lo", c_void_p), ("f_lo", c_void_p), ("s_lo", c_void_p), ("live_ctx", c_void_p), ("live_pred", c_void_p),
                ("alpha", c_void_p), ("flow", c_void_p), ("raw_output", c_void_p), ("out_full", c_void_p),
                ("norm", c_void_p), ("score", c_void_p), ("stages", C.c_int), ("storage", C.c_int)]


class DecodeBwd(C.Structure):
    _fields_ = [("f", DecodeFwd),
                ("d_output", c_void_p), ("d_raw_alpha", c_void_p), ("d_raw_output", c_void_p), ("d_flow", c_void_p), ("d_alpha", c_void_p),
                ("d_input", c_void_p), ("d_tgt_grid_obj", c_void_p), ("d_src_grid_obj", c_void_p),
                ("d_tgt_grid_bg", c_void_p), ("d_src_grid_bg", c_void_p), ("d_occ", c_void_p),
                ("d_obj_alpha", c_void_p), ("d_bg_alpha", c_void_p), ("d_cls", c_void_p),
                ("d_alpha_acc", c_void_p), ("d_f_lo", c_void_p), ("d_a_lo", c_void_p),
                ("d_prof_p", c_void_p), ("d_prof_sum", c_void_p),
                ("red_ctas", C.c_int), ("occ_part", c_void_p), ("prof_p_part", c_void_p), ("cls_part", c_void_p),
                ("up_tab", c_void_p), ("glue", c_void_p), ("stages", C.c_int),
                ("det_base", c_void_p), ("det_shadow", c_void_p), ("det_n", C.c_longlong), ("det_scale", c_void_p)]


class WifFuseFwd(C.Structure):
    _fields_ = [("B", C.c_int), ("Tc", C.c_int), ("Tp", C.c_int), ("Cr", C.c_int), ("HW", C.c_int), ("ab", C.c_int),
                ("raw_output", c_void_p), ("unet_out", c_void_p), ("frame", c_void_p)]


class WifFuseBwd(C.Structure):
    _fields_ = [("f", WifFuseFwd), ("d_frame", c_void_p), ("d_raw_output", c_void_p), ("d_unet_out", c_void_p)]


class WarpField(C.Structure):
    _fields_ = [("n", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("delta", C.c_float), ("field", c_void_p), ("grid", c_void_p), ("out", c_void_p)]


class Resize(C.Structure):
    _fields_ = [("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("H", C.c_int), ("W", C.c_int), ("in", c_void_p), ("out", c_void_p)]


class PackInput(C.Structure):
    _fields_ = [("n", C.c_int), ("Nl", C.c_int), ("HW", C.c_int), ("on", C.c_float), ("off", C.c_float),
                ("rgb_u8", c_void_p), ("rgb_f32", c_void_p), ("label", c_void_p), ("input", c_void_p), ("storage", C.c_int)]


class FramesU8(C.Structure):
    _fields_ = [("n", C.c_int), ("HW", C.c_int), ("lo", C.c_float), ("hi", C.c_float), ("frames", c_void_p), ("out", c_void_p)]


class Blur(C.Structure):
    _fields_ = [("n", C.c_int), ("H", C.c_int), ("W", C.c_int), ("ksize", C.c_int), ("sigma", C.c_float), ("in", c_void_p), ("out", c_void_p)]


class LayerEntropy(C.Structure):
    _fields_ = [("n", C.c_int), ("L", C.c_int), ("HW", C.c_int), ("alpha", c_void_p), ("entropy", c_void_p), ("fg", c_void_p)]


class LayerEntropyBwd(C.Structure):
    _fields_ = [("f", LayerEntropy), ("d_entropy", c_void_p), ("d_fg", c_void_p), ("d_alpha", c_void_p)]


class PoseDis(C.Structure):
    _fields_ = [("n", C.c_int), ("No", C.c_int), ("ho", C.c_int), ("wo", C.c_int), ("HW", C.c_int), ("eps", C.c_float),
                ("mov", c_void_p), ("fg", c_void_p), ("pose", c_void_p), ("grid", c_void_p),
                ("cell_min", c_void_p), ("center_min", c_void_p), ("cell_arg", c_void_p), ("center_arg", c_void_p)]


class PoseDisBwd(C.Structure):
    _fields_ = [("f", PoseDis), ("d_cell", c_void_p), ("d_center", c_void_p), ("d_fg", c_void_p), ("d_mov", c_void_p),
                ("ctas", C.c_int), ("part", c_void_p), ("d_pose", c_void_p)]


class ObjFlow(C.Structure):
    _fields_ = [("n", C.c_int), ("L", C.c_int), ("HW", C.c_int), ("alpha", c_void_p), ("flow", c_void_p), ("ctas", C.c_int),
                ("part", c_void_p), ("mom", c_void_p), ("dev_map", c_void_p)]


class ObjFlowBwd(C.Structure):
    _fields_ = [("f", ObjFlow), ("d_map", c_void_p), ("tsum", c_void_p), ("d_alpha", c_void_p)]


class Conv3x3(C.Structure):
    _fields_ = [("n", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Tc", C.c_int), ("Tp", C.c_int),
                ("in", c_void_p), ("weight", c_void_p), ("out", c_void_p)]


class Conv3x3Wgrad(C.Structure):
    _fields_ = [("c", Conv3x3), ("dout", c_void_p), ("ctas", C.c_int), ("part", c_void_p), ("dweight", c_void_p)]


# flags of Geom.flags (include/waldo_b200.h)
F_RESTRICT_CTX, F_FILTER, F_WEIGHT_CLS, F_HAS_CLS, F_IS_OBJ, F_INCLUDE_SELF, F_USE_DISOCC, F_OCC_PAIRS = (1 << i for i in range(8))

MAX_LAYERS, MAX_CH, MAX_LYT, MAX_TPS_K = 17, 24, 21, 256
ABI_VERSION = 2
ST_F32, ST_BF16 = 0, 1   # DecodeFwd.storage

STRUCT_OF = {"waldo_tps_fwd_t": TpsFwd, "waldo_tps_bwd_t": TpsBwd, "waldo_invwarp_fwd_t": InvWarpFwd,
             "waldo_invwarp_bwd_t": InvWarpBwd, "waldo_geom_t": Geom, "waldo_decode_fwd_t": DecodeFwd,
             "waldo_decode_bwd_t": DecodeBwd, "waldo_wif_fuse_fwd_t": WifFuseFwd, "waldo_wif_fuse_bwd_t": WifFuseBwd,
             "waldo_pack_input_t": PackInput, "waldo_warp_field_t": WarpField, "waldo_resize_t": Resize,
             "waldo_frames_u8_t": FramesU8, "waldo_blur_t": Blur, "waldo_layer_entropy_t": LayerEntropy,
             "waldo_layer_entropy_bwd_t": LayerEntropyBwd, "waldo_conv3x3_t": Conv3x3, "waldo_conv3x3_wgrad_t": Conv3x3Wgrad,
             "waldo_pose_dis_t": PoseDis, "waldo_pose_dis_bwd_t": PoseDisBwd,
             "waldo_obj_flow_t": ObjFlow, "waldo_obj_flow_bwd_t": ObjFlowBwd}

EXPORTS = ["waldo_last_error", "waldo_abi_version", "waldo_has_device_code", "waldo_launch_count", "waldo_tps_fwd", "waldo_tps_bwd",
           "waldo_invwarp_fwd", "waldo_invwarp_bwd", "waldo_occ_fwd", "waldo_occ_bwd", "waldo_decode_fwd",
           "waldo_decode_bwd", "waldo_wif_fuse_fwd", "waldo_wif_fuse_bwd", "waldo_pack_input", "waldo_warp_field_fwd", "waldo_resize_bilinear_fwd",
           "waldo_frames_to_u8", "waldo_blur_fwd", "waldo_blur_bwd", "waldo_layer_entropy_fwd", "waldo_layer_entropy_bwd", "waldo_conv3x3_fwd", "waldo_conv3x3_wgrad",
           "waldo_pose_dis_fwd", "waldo_pose_dis_bwd", "waldo_obj_flow_fwd", "waldo_obj_flow_bwd"]

_lock = threading.Lock()
_lib = None
_allow_host_pointers = False   # flipped only by tests/emu (kernel-logic emulation), never by the package


def _declare(lib):
    lib.waldo_last_error.restype = C.c_char_p
    lib.waldo_abi_version.restype = C.c_int
    lib.waldo_has_device_code.restype = C.c_int
    lib.waldo_launch_count.restype = C.c_longlong
    for name, st in (("waldo_tps_fwd", TpsFwd), ("waldo_tps_bwd", TpsBwd), ("waldo_invwarp_fwd", InvWarpFwd),
                     ("waldo_invwarp_bwd", InvWarpBwd), ("waldo_decode_fwd", DecodeFwd), ("waldo_decode_bwd", DecodeBwd),
                     ("waldo_wif_fuse_fwd", WifFuseFwd), ("waldo_wif_fuse_bwd", WifFuseBwd), ("waldo_pack_input", PackInput), ("waldo_warp_field_fwd", WarpField),
                     ("waldo_resize_bilinear_fwd", Resize), ("waldo_frames_to_u8", FramesU8), ("waldo_blur_fwd", Blur), ("waldo_blur_bwd", Blur),
                     ("waldo_layer_entropy_fwd", LayerEntropy), ("waldo_layer_entropy_bwd", LayerEntropyBwd), ("waldo_conv3x3_fwd", Conv3x3), ("waldo_conv3x3_wgrad", Conv3x3Wgrad),
                     ("waldo_pose_dis_fwd", PoseDis), ("waldo_pose_dis_bwd", PoseDisBwd),
                     ("waldo_obj_flow_fwd", ObjFlow), ("waldo_obj_flow_bwd", ObjFlowBwd)):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(st), c_void_p]
        fn.restype = C.c_int
    lib.waldo_occ_fwd.argtypes = [C.c_int, C.c_int, c_void_p, c_void_p, c_void_p]
    lib.waldo_occ_fwd.restype = C.c_int
    lib.waldo_occ_bwd.argtypes = [C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.waldo_occ_bwd.restype = C.c_int
    return lib


def load(build_if_missing: bool = True):
    """Load (building first if the in-tree .so is missing or stale and nvcc is present) the sm_100a library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing and not os.environ.get("WALDO_B200_LIB"):
            from . import build as _build
            try:
                if _build.stale():
                    _build.build()
            except Exception as e:  # no nvcc on this box: fine as long as a prebuilt library travels with the tree
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(f"waldo_b200: CUDA library missing and cannot be built: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"waldo_b200: {LIB_PATH} not found; run `python -m waldo_b200.build` (needs nvcc). "
                               "There is no CPU fallback.")
        lib = _declare(C.CDLL(LIB_PATH))
        if lib.waldo_abi_version() != ABI_VERSION:
            raise RuntimeError("waldo_b200: ABI version mismatch between _lib.py and libwaldo_b200.so")
        if lib.waldo_has_device_code() != 1:
            raise RuntimeError("waldo_b200: library was built without device code; refusing to use it")
        _lib = lib
        return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"waldo_b200.{what} failed ({rc}): {load().waldo_last_error().decode()}")


def device_of(dev):
    """Device guard for a library call: kernels launch on the calling thread's CURRENT device, while the stream handed over
    is the current stream of the TENSORS' device (reference precedent: the OptionalCUDAGuard of bias_act.cpp:54)."""
    import contextlib
    dev = torch.device(dev)
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def call(fn, arg, ref: torch.Tensor, what: str):
    """fn(&arg, current stream of ref's device) under a device guard; raises on a non-zero return code."""
    with device_of(ref.device):
        check(fn(C.byref(arg), stream_of(ref)), what)


def stream_of(t: torch.Tensor):
    if t.is_cuda:
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return C.c_void_p(0)


def ptr(t, dtype=torch.float32, name="tensor"):
    """Raw pointer of a tensor that must be dense, of `dtype` and on a CUDA device (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda and not _allow_host_pointers:
        raise RuntimeError(f"waldo_b200: {name} must be a CUDA tensor (got {t.device}); there is no CPU path")
    if t.dtype != dtype:
        raise RuntimeError(f"waldo_b200: {name} must be {dtype} (got {t.dtype})")
    if not t.is_contiguous():
        raise RuntimeError(f"waldo_b200: {name} must be contiguous")
    return C.c_void_p(t.data_ptr())
