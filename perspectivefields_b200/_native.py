"""ctypes binding of libpf_b200.so (C ABI declared in include/pf_b200.h) and its in-tree build recipe.

There is no CPU fallback: importing this module works without a GPU (so the ABI can be inspected), but every
compute entry point needs the CUDA library and a B200.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libpf_b200.so")
SRC_DIR = os.path.join(_HERE, "csrc")
HEADER = os.path.join(ROOT, "include", "pf_b200.h")

PF_F32, PF_BF16 = 0, 1
PF_PARAM_NONE, PF_PARAM_CENTERED, PF_PARAM_UNCENTERED = 0, 1, 2

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


class pf_model_desc(ctypes.Structure):
    _fields_ = [("gravity_classes", ctypes.c_int), ("latitude_classes", ctypes.c_int), ("param_net", ctypes.c_int),
                ("param_input_size", ctypes.c_int), ("pixel_mean", ctypes.c_float * 3), ("pixel_std", ctypes.c_float * 3)]


class pf_camera(ctypes.Structure):
    """include/pf_b200.h: struct pf_camera."""
    _fields_ = [("height", ctypes.c_int32), ("width", ctypes.c_int32), ("focal_rel", ctypes.c_double), ("elevation", ctypes.c_double),
                ("roll", ctypes.c_double), ("cx_rel", ctypes.c_double), ("cy_rel", ctypes.c_double),
                ("up_offset", ctypes.c_int64), ("lat_offset", ctypes.c_int64)]


class pf_batch(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int),
                ("images_u8", ctypes.c_void_p), ("image_offset", ctypes.POINTER(ctypes.c_int64)),
                ("images_chw", ctypes.c_void_p),
                ("height", ctypes.POINTER(ctypes.c_int32)), ("width", ctypes.POINTER(ctypes.c_int32)),
                ("pred_gravity", ctypes.c_void_p), ("pred_latitude", ctypes.c_void_p),
                ("gravity_original", ctypes.c_void_p), ("gravity_original_offset", ctypes.POINTER(ctypes.c_int64)),
                ("latitude_original", ctypes.c_void_p), ("latitude_original_offset", ctypes.POINTER(ctypes.c_int64)),
                ("params", ctypes.c_void_p)]


def _sources():
    return sorted(os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".cu", ".cuh"))) + [HEADER]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    """Compile csrc/pf_b200.cu for sm_100a into libpf_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(SRC_DIR, "pf_b200.cu"), "-o", LIB_PATH + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load libpf_b200.so (raises if it has not been built: the product has no other compute path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(perspectivefields_b200 has no CPU or PyTorch fallback)")
    # PF_B200_LIB: an alternative build of the SAME library (A/B timing of kernel changes on one box: tools/ab.sh)
    L = ctypes.CDLL(os.environ.get("PF_B200_LIB") or LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    sig = {
        "pf_abi_version": (i32, []),
        "pf_last_error": (ctypes.c_char_p, []),
        "pf_kernel_launch_count": (i64, []),
        "pf_create": (i32, [i32, ctypes.POINTER(pf_model_desc), ctypes.POINTER(vp)]),
        "pf_destroy": (i32, [vp]),
        "pf_set_weight": (i32, [vp, ctypes.c_char_p, vp, i64, i32]),
        "pf_finalize": (i32, [vp]),
        "pf_workspace_bytes": (i64, [vp, i32, i32]),
        "pf_forward": (i32, [vp, ctypes.POINTER(pf_batch), vp, i64, vp]),
        "pf_profile_enable": (i32, [vp, i32]),
        "pf_profile_read": (i32, [vp, ctypes.POINTER(ctypes.c_double)]),
        "pf_profile_kernels_enable": (i32, [vp, i32]),
        "pf_profile_kernels_read": (i32, [vp, ctypes.c_char_p, i32]),
        "pf_debug_enable": (i32, [vp, i32]),
        "pf_debug_count": (i32, [vp]),
        "pf_debug_name": (ctypes.c_char_p, [vp, i32]),
        "pf_debug_numel": (i64, [vp, ctypes.c_char_p]),
        "pf_debug_copy": (i32, [vp, ctypes.c_char_p, vp, i64, vp]),
        "pf_op_conv_gemm": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp, vp]),
        "pf_set_option": (i32, [vp, ctypes.c_char_p, i32]),
        "pf_camera_fields": (i32, [i32, ctypes.POINTER(pf_camera), i32, vp, vp, vp]),
        "pf_op_layernorm": (i32, [vp, vp, i64, i32, vp, vp, f32, vp]),
        "pf_op_attention": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "pf_op_attention_mma": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "pf_op_attention_tc": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "pf_op_dwconv3x3_gelu": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "pf_op_dwconv7x7": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "pf_op_upsample2x": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "pf_op_preprocess": (i32, [vp, i32, i32, ctypes.POINTER(f32), ctypes.POINTER(f32), vp, vp]),
        "pf_op_resize_u8": (i32, [vp, i32, i32, i32, i32, vp, vp]),
        "pf_op_resize_f32": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
        "pf_op_argmax_decode": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "pf_op_fill_stream": (i32, [vp, i64, f32, vp]),
        "pf_op_pred_argmax_decode": (i32, [vp, i32, i32, vp, vp, vp, i32, i32, i32, i32, vp]),
        "pf_op_postprocess": (i32, [vp, vp, i32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), vp, ctypes.POINTER(i64), vp,
                                    ctypes.POINTER(i64), i32, vp]),
        "pf_comm_unique_id": (i32, [vp]),
        "pf_comm_create": (i32, [i32, i32, i32, vp, ctypes.POINTER(vp)]),
        "pf_comm_destroy": (i32, [vp]),
        "pf_gather": (i32, [vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int32), vp]),
        "pf_jpeg_create": (i32, [i32, i32, ctypes.POINTER(vp)]),
        "pf_jpeg_destroy": (i32, [vp]),
        "pf_jpeg_info": (i32, [vp, vp, i64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
        "pf_jpeg_decode_batch": (i32, [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int32),
                                       ctypes.POINTER(ctypes.c_int32), vp, ctypes.POINTER(i64), vp]),
    }
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        except AttributeError:
            if os.environ.get("PF_B200_LIB"):   # an older build under A/B test may predate an entry point
                continue
            raise
        fn.restype, fn.argtypes = res, args
    if L.pf_abi_version() != 2:
        raise RuntimeError("libpf_b200.so ABI version mismatch")
    _lib = L
    return L


EXPORTS = ["pf_abi_version", "pf_last_error", "pf_kernel_launch_count", "pf_create", "pf_destroy", "pf_set_weight",
           "pf_finalize", "pf_workspace_bytes", "pf_forward", "pf_profile_enable", "pf_profile_read", "pf_profile_kernels_enable",
           "pf_profile_kernels_read", "pf_set_option", "pf_debug_enable", "pf_debug_count", "pf_debug_name", "pf_debug_numel",
           "pf_debug_copy", "pf_camera_fields", "pf_comm_unique_id", "pf_comm_create", "pf_comm_destroy", "pf_gather",
           "pf_jpeg_create", "pf_jpeg_destroy", "pf_jpeg_info", "pf_jpeg_decode_batch",
           "pf_op_conv_gemm", "pf_op_layernorm", "pf_op_attention", "pf_op_attention_mma", "pf_op_attention_tc", "pf_op_dwconv3x3_gelu", "pf_op_dwconv7x7",
           "pf_op_upsample2x", "pf_op_preprocess", "pf_op_fill_stream", "pf_op_resize_u8", "pf_op_resize_f32", "pf_op_argmax_decode",
           "pf_op_pred_argmax_decode", "pf_op_postprocess"]


class PfError(RuntimeError):
    pass


def check(status):
    if status < 0:
        raise PfError(f"libpf_b200: {lib().pf_last_error().decode()} (status {status})")
    return status
