"""ctypes binding of include/rcvvote.h.  The library is required: if librcvvote.so is missing or no
sm_100 GPU is present every entry point raises -- there is deliberately no CPU or PyTorch fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCV_LIB_PATH", os.path.join(HERE, "librcvvote.so"))   # override: kernel-variant experiments only

RCV_ABI_VERSION = 2
RCV_OK = 0
RCV_F32, RCV_F64, RCV_U16 = 0, 1, 2
RCV_POLICY_LM, RCV_POLICY_YCBGEN = 0, 1
RCV_MASK_RADIUS_NONZERO, RCV_MASK_RADIUS_POSITIVE, RCV_MASK_SEM_GT, RCV_MASK_SEM_GE, RCV_MASK_MAX_RADIUS = 1, 2, 4, 8, 16
RCV_ST_OK, RCV_ST_EMPTY_MASK, RCV_ST_BAD_GRID, RCV_ST_D_EXCEEDS_CAP, RCV_ST_POINT_OVERFLOW, RCV_ST_UNIT_OVERFLOW = 0, 1, 2, 4, 8, 16
RCV_ST_VOLUME_SKIPPED = 32

EXPORTS = ["rcv_create", "rcv_destroy", "rcv_last_error", "rcv_abi_version", "rcv_backproject", "rcv_vote_points", "rcv_vote_frames",
           "rcv_vote_frames_host", "rcv_argmax_volume", "rcv_head_1x1", "rcv_horn_batch", "rcv_horn_batch_host", "rcv_launch_count", "rcv_last_h2d_bytes",
           "rcv_last_vote_kernel_ms", "rcv_vote_kernel_times", "rcv_ubench_smem_atomics", "rcv_add_metric_batch", "rcv_scene_clouds", "rcv_scene_clouds_last", "rcv_icp_batch", "rcv_head_vote_frames",
           "rcv_conv7_head", "rcv_conv7_head_vote_frames"]


class rcv_config(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("max_items", C.c_int), ("max_points_total", C.c_longlong), ("max_grid", C.c_int),
                ("max_units", C.c_int), ("image_pixels", C.c_longlong), ("max_model_points", C.c_int), ("head_items", C.c_int)]


class rcv_vote_params(C.Structure):
    _fields_ = [("acc_unit", C.c_double), ("radius_scale", C.c_double), ("grid_policy", C.c_int), ("radius_dtype", C.c_int)]


class rcv_frame_params(C.Structure):
    _fields_ = [("height", C.c_int), ("width", C.c_int), ("depth_dtype", C.c_int), ("depth_div", C.c_double), ("xyz_div", C.c_double),
                ("mask_flags", C.c_int), ("sem_threshold", C.c_float), ("k_stride", C.c_int), ("max_radii_stride", C.c_int)]


_lib = None


def load():
    """Load librcvvote.so (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build
    try:
        if "RCV_LIB_PATH" not in os.environ and build.needs_build():
            build.build_library()
    except Exception as e:  # no nvcc on the box: fall through to the prebuilt library if there is one
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("librcvvote.so is missing and cannot be built: %s" % e)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("librcvvote.so not found at %s; run `python -m rcvpose_b200.build`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ip, llp = C.c_void_p, C.c_void_p, C.c_void_p
    L.rcv_abi_version.restype = C.c_int
    L.rcv_create.restype = C.c_int
    L.rcv_create.argtypes = [C.c_int, C.POINTER(rcv_config), C.POINTER(vp)]
    L.rcv_destroy.restype = None
    L.rcv_destroy.argtypes = [vp]
    L.rcv_last_error.restype = C.c_char_p
    L.rcv_last_error.argtypes = [vp]
    L.rcv_backproject.restype = C.c_int
    L.rcv_backproject.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_longlong, ip, vp]
    L.rcv_vote_points.restype = C.c_int
    L.rcv_vote_points.argtypes = [vp, vp, vp, llp, C.c_int, C.POINTER(rcv_vote_params), vp, ip, llp, ip, ip, ip, vp, C.c_longlong, vp]
    L.rcv_vote_frames.restype = C.c_int
    L.rcv_vote_frames.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(rcv_frame_params), C.POINTER(rcv_vote_params),
                                  vp, ip, llp, ip, ip, ip, vp]
    L.rcv_vote_frames_host.restype = C.c_int
    L.rcv_vote_frames_host.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(rcv_frame_params), C.POINTER(rcv_vote_params),
                                       vp, ip, llp, ip, ip, ip, C.c_int, vp]
    L.rcv_argmax_volume.restype = C.c_int
    L.rcv_argmax_volume.argtypes = [vp, vp, C.c_int, ip, ip, vp]
    L.rcv_horn_batch.restype = C.c_int
    L.rcv_horn_batch.argtypes = [vp, vp, C.c_longlong, vp, C.c_int, C.c_int, vp, vp]
    L.rcv_add_metric_batch.restype = C.c_int
    L.rcv_add_metric_batch.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp]
    L.rcv_scene_clouds_last.restype = C.c_int
    L.rcv_scene_clouds_last.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.POINTER(rcv_frame_params), C.c_double, vp, C.c_longlong, vp, vp, vp]
    L.rcv_scene_clouds.restype = C.c_int
    L.rcv_scene_clouds.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(rcv_frame_params), C.c_double, vp, C.c_longlong, llp,
                                   ip, vp]
    L.rcv_icp_batch.restype = C.c_int
    L.rcv_icp_batch.argtypes = [vp, vp, C.c_int, vp, llp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp, ip, vp]
    L.rcv_horn_batch_host.restype = C.c_int
    L.rcv_horn_batch_host.argtypes = [vp, vp, C.c_longlong, vp, C.c_int, C.c_int, vp, vp]
    L.rcv_head_1x1.restype = C.c_int
    L.rcv_head_1x1.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_longlong, vp]
    L.rcv_head_vote_frames.restype = C.c_int
    L.rcv_head_vote_frames.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(rcv_frame_params), C.POINTER(rcv_vote_params),
                                       vp, ip, llp, ip, ip, ip, vp, vp]
    L.rcv_conv7_head.restype = C.c_int
    L.rcv_conv7_head.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.rcv_conv7_head_vote_frames.restype = C.c_int
    L.rcv_conv7_head_vote_frames.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_void_p), vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(rcv_frame_params),
                                             C.POINTER(rcv_vote_params), vp, ip, llp, ip, ip, ip, vp, vp]
    L.rcv_launch_count.restype = C.c_longlong
    L.rcv_launch_count.argtypes = [vp]
    L.rcv_last_h2d_bytes.restype = C.c_longlong
    L.rcv_last_h2d_bytes.argtypes = [vp]
    L.rcv_last_vote_kernel_ms.restype = C.c_float
    L.rcv_last_vote_kernel_ms.argtypes = [vp]
    L.rcv_vote_kernel_times.restype = C.c_int
    L.rcv_vote_kernel_times.argtypes = [vp, vp, C.c_int]
    L.rcv_ubench_smem_atomics.restype = C.c_int
    L.rcv_ubench_smem_atomics.argtypes = [vp, C.POINTER(C.c_double)]
    if L.rcv_abi_version() != RCV_ABI_VERSION:
        raise RuntimeError("librcvvote.so ABI version mismatch")
    _lib = L
    return L
