"""ctypes binding of libsgam_b200.so (include/sgam_b200.h).

There is no CPU fallback: if the shared library is missing the import of any product op raises, and every
op checks its status code and raises RuntimeError(sgam_last_error()).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgam_b200.so")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_sz = ctypes.c_size_t
c_u32 = ctypes.c_uint32

# name -> (restype, argtypes); mirrors include/sgam_b200.h one to one (tests/test_abi.py checks the header)
SIGNATURES = {
    "sgam_last_error": (ctypes.c_char_p, []),
    "sgam_version": (c_i, []),
    "sgam_sm_count": (c_i, [c_i]),
    "sgam_launch_count": (ctypes.c_ulonglong, []),
    "sgam_splat_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "sgam_splat_forward": (c_i, [c_p, c_ll, c_ll, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i,
                                 c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "sgam_median_blur3": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p]),
    "sgam_depth_code": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "sgam_inverse_warp": (c_i, [c_p, c_ll, c_ll, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "sgam_frame_outputs": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "sgam_vq_workspace_bytes": (c_sz, [c_i]),
    "sgam_vq_nearest": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "sgam_vq_topk_workspace_bytes": (c_sz, [c_i, c_i]),
    "sgam_vq_topk_sample": (c_i, [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, ctypes.c_ulonglong, c_i,
                                  c_p, c_p, c_p, c_p, c_p, c_p]),
    "sgam_vq_norms": (c_i, [c_p, c_p, c_i, c_i, c_p]),
    "sgam_vq_tc_workspace_bytes": (c_sz, [c_i, c_i]),
    "sgam_vq_nearest_tc": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_f, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "sgam_stem_conv": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "sgam_conv2d": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sgam_gn_splits": (c_i, [c_ll]),
    "sgam_groupnorm": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_i, c_i, c_p]),
    "sgam_gemm_nt": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_ll, c_ll, c_ll, c_f, c_p]),
    "sgam_softmax_rows": (c_i, [c_p, c_ll, c_i, c_p]),
    "sgam_stem_conv_split": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "sgam_split_bf16": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sgam_groupnorm_split": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_i, c_i, c_p]),
    "sgam_softmax_split": (c_i, [c_p, c_p, c_p, c_ll, c_i, c_p]),
    "sgam_tc_supported_conv": (c_i, [c_i, c_i, c_i, c_i, c_i, c_i]),
    "sgam_conv2d_tc": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p,
                             ctypes.POINTER(c_i), c_p]),
    "sgam_groupnorm_split_apply": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_i, c_i, c_p]),
    "sgam_conv2d_tc_splitk_floats": (c_ll, [c_i, c_i, c_i, c_i, c_i, c_i, c_i]),
    "sgam_tc_gn_partial_floats": (c_ll, [c_i, c_i, c_i]),
    "sgam_groupnorm_split_fused": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sgam_gemm_nt_tc": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_i, c_p]),
    "sgam_conv2d_tc_up2_supported": (c_i, [c_i, c_i, c_i, c_i, c_i]),
    "sgam_conv2d_tc_up2": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "sgam_stem_conv_in": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sgam_gn_head_conv": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sgam_attention_tc_supported": (c_i, [c_i, c_i, c_i]),
    "sgam_attention_tc_splits": (c_i, [c_i, c_i]),
    "sgam_attention_tc_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "sgam_attention_tc": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_i, c_p, c_i, c_p]),
    "sgam_gn_conv2d_tc_supported": (c_i, [c_i, c_i, c_i, c_i, c_i]),
    "sgam_gn_conv2d_tc": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "sgam_qkv_tc_supported": (c_i, [c_i, c_i, c_i, c_i]),
    "sgam_qkv_tc": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sgam_unproject_points": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    "sgam_tsdf_volume_bytes": (c_sz, [c_i, c_i, c_i, c_i]),
    "sgam_tsdf_block_bytes": (c_sz, [c_i]),
    "sgam_tsdf_integrate": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f,
                                  c_p, c_u32, c_p, c_p, c_p, c_p, c_p, c_p]),
    "sgam_tsdf_raycast": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_f, c_i, c_i, c_f, c_f, c_f,
                                c_p, c_p]),
    "sgam_tsdf_extract": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p, c_p]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m sgam_neurips22_b200.build` "
                "(there is no CPU fallback for the SGAM hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError = header / library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = load().sgam_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {status}: {msg}")
