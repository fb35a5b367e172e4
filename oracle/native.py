"""ctypes bindings for oracle/csrc/oracle.c.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "oracle.c")
_SRCS = [_SRC, os.path.join(_HERE, "csrc", "tsdf_oracle.c")]
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "liboracle.so")
_lib = None

DATASET_ID = {"clevr-infinite": 0, "google_earth": 1}


def build(force=False):
    """gcc recipe for the C restatement (also run by __graft_entry__.build())."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in _SRCS):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-mfma", "-fopenmp",
               "-o", _SO, *_SRCS, "-lm"]
        subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def splat_forward(src_rgb, src_depth, K_tgt, Kinv_src, T, zmin=False):
    """src_rgb [B,N,3,H,W], src_depth [B,N,H,W], K_tgt [B,3,3], Kinv_src [B,N,3,3], T [B,N,4,4] (numpy fp32)."""
    src_rgb, src_depth, K_tgt, Kinv_src, T = map(_f, (src_rgb, src_depth, K_tgt, Kinv_src, T))
    B, N, H, W = src_depth.shape
    out = dict(proj_rgb=np.empty((B, 3, H, W), np.float32), proj_depth=np.empty((B, 1, H, W), np.float32),
               merge_rgb=np.empty((B, 3, H, W), np.float32), merge_depth=np.empty((B, 1, H, W), np.float32),
               mask=np.empty((B, 1, H, W), np.uint8), winner=np.empty((B, H, W), np.int32),
               inbounds=np.empty((B, H * W * N), np.uint8))
    lib().oracle_splat_forward(_p(src_rgb), _p(src_depth), _p(K_tgt), _p(Kinv_src), _p(T),
                               B, N, H, W, int(bool(zmin)), _p(out["proj_rgb"]), _p(out["proj_depth"]),
                               _p(out["merge_rgb"]), _p(out["merge_depth"]), _p(out["mask"]),
                               _p(out["winner"]), _p(out["inbounds"]))
    return out


def median_blur3(x):
    x = _f(x)
    out = np.empty_like(x)
    H, W = x.shape[-2:]
    lib().oracle_median_blur3(_p(x), _p(out), int(x.size // (H * W)), H, W)
    return out


def depth_code(depth, mask, dataset):
    depth = _f(depth)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty_like(depth)
    lib().oracle_depth_code(_p(depth), _p(mask), _p(out), ctypes.c_size_t(depth.size), DATASET_ID[dataset])
    return out


def depth_decode(code, dataset):
    code = _f(code)
    out = np.empty_like(code)
    lib().oracle_depth_decode(_p(code), _p(out), ctypes.c_size_t(code.size), DATASET_ID[dataset])
    return out


def pack_u8(rgb):
    rgb = _f(rgb)
    _, H, W = rgb.shape
    out = np.empty((H, W, 3), np.uint8)
    lib().oracle_pack_u8(_p(rgb), _p(out), H, W)
    return out


def inverse_warp(src_rgb, src_depth, tgt_depth, Kinv_tgt, proj):
    src_rgb, src_depth, tgt_depth, Kinv_tgt, proj = map(_f, (src_rgb, src_depth, tgt_depth, Kinv_tgt, proj))
    B, N, H, W = src_depth.shape
    out = np.empty((B, 3, H, W), np.float32)
    best = np.empty((B, H, W), np.int32)
    lib().oracle_inverse_warp(_p(src_rgb), _p(src_depth), _p(tgt_depth), _p(Kinv_tgt), _p(proj),
                              B, N, H, W, _p(out), _p(best))
    return out, best


def vq_nearest(z_tokens, codebook):
    """z_tokens [T,D], codebook [n_e,D] -> (idx int64 [T], dmin [T], second [T])."""
    z, E = _f(z_tokens), _f(codebook)
    T, D = z.shape
    idx = np.empty(T, np.int64)
    dmin = np.empty(T, np.float32)
    d2 = np.empty(T, np.float32)
    lib().oracle_vq_nearest(_p(z), _p(E), T, E.shape[0], D, _p(idx), _p(dmin), _p(d2))
    return idx, dmin, d2


def unproject_world(depth, Kinv, Rt_inv):
    depth = _f(depth)
    H, W = depth.shape
    Kinv = np.ascontiguousarray(Kinv, np.float64)
    Rt_inv = np.ascontiguousarray(Rt_inv, np.float64)
    out = np.empty((H * W, 3), np.float64)
    lib().oracle_unproject_world(_p(depth), _p(Kinv), _p(Rt_inv), H, W, _p(out))
    return out


# ---------------------------------------------------------------------------------------------- TSDF (parity unpinned)
class TsdfGrid(ctypes.Structure):
    _fields_ = [("ox", ctypes.c_int), ("oy", ctypes.c_int), ("oz", ctypes.c_int),
                ("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int),
                ("voxel_length", ctypes.c_float), ("sdf_trunc", ctypes.c_float)]


class TsdfVolume:
    """Host twin of sgam_neurips22_b200.tsdf.TSDFVolume on oracle/csrc/tsdf_oracle.c (same layout, same arguments)."""

    def __init__(self, voxel_length, sdf_trunc, origin_units, dims_units, with_color=True):
        self.grid = TsdfGrid(*[int(v) for v in origin_units], *[int(v) for v in dims_units], float(voxel_length), float(sdf_trunc))
        n = int(np.prod(dims_units))
        self.stamp = np.zeros(n, np.uint32)
        self.vol = np.zeros((n, 4096, 2), np.float32)
        self.color = np.zeros((n, 4096, 3), np.float32) if with_color else None
        self.frame = 0

    def integrate(self, depth, rgb, K4, world2cam, stride=4, depth_trunc=20.0):
        """depth [H,W] fp32, rgb [H,W,3] fp32 in [-1,1] or None, K4 = (fx,fy,cx,cy), world2cam 4x4 float64."""
        depth = _f(depth)
        H, W = depth.shape
        rgb = _f(rgb) if (rgb is not None and self.color is not None) else None
        w2c = np.asarray(world2cam, np.float64)
        c2w = np.ascontiguousarray(np.linalg.inv(w2c)[:3], np.float64)
        w2c32 = np.ascontiguousarray(w2c[:3], np.float32)
        K = np.ascontiguousarray(K4, np.float64)
        self.frame += 1
        L = lib()
        L.oracle_tsdf_touch(_p(depth), H, W, _p(c2w), _p(K), int(stride), ctypes.c_float(depth_trunc),
                            ctypes.byref(self.grid), _p(self.stamp), ctypes.c_uint32(self.frame))
        L.oracle_tsdf_integrate(_p(depth), _p(rgb), H, W, _p(w2c32), _p(K), ctypes.c_float(depth_trunc),
                                ctypes.byref(self.grid), _p(self.stamp), ctypes.c_uint32(self.frame), _p(self.vol),
                                _p(self.color) if rgb is not None else None)

    def render_depth(self, K4, world2cam, H, W, pixel_center=0.5, z_near=0.05, z_far=20.0, step_vox=0.5):
        c2w32 = np.ascontiguousarray(np.linalg.inv(np.asarray(world2cam, np.float64))[:3], np.float32)
        K = np.ascontiguousarray(K4, np.float64)
        out = np.empty((H, W), np.float32)
        lib().oracle_tsdf_raycast(ctypes.byref(self.grid), _p(self.stamp), _p(self.vol), _p(c2w32), _p(K),
                                  ctypes.c_float(pixel_center), H, W, ctypes.c_float(z_near), ctypes.c_float(z_far),
                                  ctypes.c_float(step_vox), _p(out))
        return out

    def extract_point_cloud(self):
        L = lib()
        L.oracle_tsdf_extract.restype = ctypes.c_long
        col = _p(self.color) if self.color is not None else None
        n = L.oracle_tsdf_extract(ctypes.byref(self.grid), _p(self.stamp), _p(self.vol), col, None, None)
        xyz, rgb = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        L.oracle_tsdf_extract(ctypes.byref(self.grid), _p(self.stamp), _p(self.vol), col, _p(xyz), _p(rgb))
        return xyz, rgb
