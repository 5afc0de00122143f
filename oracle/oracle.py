"""ctypes/numpy wrapper of the CPU oracle (oracle/lonlat_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
reference-fallback legs.  Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = np.float32
_pf = ctypes.POINTER(ctypes.c_float)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liblonlat_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liblonlat_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.ogs_oracle_preprocess_fwd.restype = ctypes.c_int64
        _LIB.ogs_oracle_pinhole_preprocess_fwd.restype = ctypes.c_int64
        _LIB.ogs_oracle_higher_msb.restype = ctypes.c_uint32
    return _LIB


def set_seam_wrap(on):
    """Switch the oracle to the non-parity seam wrap-around extension (default off = reference behaviour)."""
    lib().ogs_oracle_set_seam_wrap(1 if on else 0)


def num_threads():
    return int(lib().ogs_oracle_num_threads())


def set_num_threads(n):
    """OpenMP threads of the following oracle calls (bench.py's single-thread CPU row)."""
    lib().ogs_oracle_set_num_threads(int(n))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype=_f):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def forward(means3D, opacities, viewmatrix, campos, W, H, background, shs=None, degree=0, colors_precomp=None,
            scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0, want_binning=True, pinhole=None):
    """Reference-order forward.  Returns a dict with every intermediate the reference keeps.
    pinhole = dict(projmatrix=[4,4], tan_fovx=, tan_fovy=, render_depth=False) selects camera_type 1."""
    L = lib()
    means3D = _c(means3D); P = means3D.shape[0]
    opacities = _c(opacities).reshape(-1); viewmatrix = _c(viewmatrix).reshape(-1); campos = _c(campos).reshape(-1)
    background = _c(background).reshape(-1)
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(_c, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    M = 0 if shs is None else shs.shape[1]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o = dict(
        radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), _f), depths=np.zeros(P, _f),
        cov3D=np.zeros((P, 6), _f), rgb=np.zeros((P, 3), _f), conic_opacity=np.zeros((P, 4), _f),
        tiles_touched=np.zeros(P, np.uint32), point_offsets=np.zeros(P, np.uint32), clamped=np.zeros((P, 3), np.uint8))
    R = 0
    if P and pinhole is not None:
        proj = _c(pinhole["projmatrix"]).reshape(-1)
        R = int(L.ogs_oracle_pinhole_preprocess_fwd(
            P, int(degree), M, _p(means3D), _p(scales), ctypes.c_float(scale_modifier), _p(rotations), _p(opacities),
            _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(viewmatrix), _p(proj), _p(campos), W, H,
            ctypes.c_float(pinhole["tan_fovx"]), ctypes.c_float(pinhole["tan_fovy"]),
            1 if pinhole.get("render_depth") else 0,
            _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]), _p(o["conic_opacity"]),
            _p(o["tiles_touched"]), _p(o["point_offsets"]), _p(o["clamped"])))
    elif P:
        R = int(L.ogs_oracle_preprocess_fwd(
            P, int(degree), M, _p(means3D), _p(scales), ctypes.c_float(scale_modifier), _p(rotations), _p(opacities),
            _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(viewmatrix), _p(campos), W, H,
            _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]), _p(o["conic_opacity"]),
            _p(o["tiles_touched"]), _p(o["point_offsets"]), _p(o["clamped"])))
    o["num_rendered"] = R
    if cov3D_precomp is not None:
        o["cov3D"] = cov3D_precomp.copy()
    if not want_binning:
        return o
    o.update(keys_unsorted=np.zeros(R, np.uint64), values_unsorted=np.zeros(R, np.uint32),
             point_list_keys=np.zeros(R, np.uint64), point_list=np.zeros(R, np.uint32),
             ranges=np.zeros((gx * gy, 2), np.uint32))
    L.ogs_oracle_bin(P, W, H, _p(o["means2D"]), _p(o["depths"]), _p(o["radii"]), _p(o["point_offsets"]),
                     ctypes.c_int64(R), _p(o["keys_unsorted"]), _p(o["values_unsorted"]), _p(o["point_list_keys"]),
                     _p(o["point_list"]), _p(o["ranges"]))
    colors = colors_precomp if colors_precomp is not None else o["rgb"]
    if pinhole is not None and pinhole.get("render_depth"):
        colors = o["rgb"]           # renderDepthCUDA blends the depth whatever the colour source
    o.update(accum_alpha=np.zeros(H * W, _f), n_contrib=np.zeros(H * W, np.uint32), out_color=np.zeros((3, H, W), _f))
    L.ogs_oracle_render_fwd(W, H, _p(o["ranges"]), _p(o["point_list"]), _p(o["means2D"]), _p(colors),
                            _p(o["conic_opacity"]), _p(background), _p(o["accum_alpha"]), _p(o["n_contrib"]),
                            _p(o["out_color"]))
    return o


def backward(fwd, dL_dout_color, means3D, viewmatrix, campos, W, H, background, shs=None, degree=0,
             colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0, pinhole=None):
    """Reference-order backward from a forward() dict.  Returns the 8 returned gradients + dL_dconic."""
    L = lib()
    means3D = _c(means3D); P = means3D.shape[0]
    viewmatrix = _c(viewmatrix).reshape(-1); campos = _c(campos).reshape(-1); background = _c(background).reshape(-1)
    shs, colors_precomp, scales, rotations = map(_c, (shs, colors_precomp, scales, rotations))
    dL = _c(dL_dout_color)
    M = 0 if shs is None else shs.shape[1]
    colors = colors_precomp if colors_precomp is not None else fwd["rgb"]
    g = dict(dL_dmeans2D=np.zeros((P, 3), _f), dL_dconic=np.zeros((P, 4), _f), dL_dopacity=np.zeros((P, 1), _f),
             dL_dcolors=np.zeros((P, 3), _f), dL_dmeans3D=np.zeros((P, 3), _f), dL_dcov3D=np.zeros((P, 6), _f),
             dL_dsh=np.zeros((P, M, 3), _f), dL_dscales=np.zeros((P, 3), _f), dL_drotations=np.zeros((P, 4), _f),
             dpx_dt=np.zeros((P, 3), _f), dpy_dt=np.zeros((P, 3), _f))
    if P == 0:
        return g
    L.ogs_oracle_render_bwd(W, H, _p(fwd["ranges"]), _p(fwd["point_list"]), _p(background), _p(fwd["means2D"]),
                            _p(fwd["conic_opacity"]), _p(colors), _p(fwd["accum_alpha"]), _p(fwd["n_contrib"]), _p(dL),
                            _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]))
    if pinhole is not None:
        proj = _c(pinhole["projmatrix"]).reshape(-1)
        L.ogs_oracle_pinhole_preprocess_bwd(
            P, int(degree), M, _p(means3D), _p(fwd["radii"]), _p(shs), _p(fwd["clamped"]), _p(scales), _p(rotations),
            ctypes.c_float(scale_modifier), _p(_c(fwd["cov3D"])), _p(viewmatrix), _p(proj), W, H,
            ctypes.c_float(pinhole["tan_fovx"]), ctypes.c_float(pinhole["tan_fovy"]), _p(campos),
            _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
            _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
        return g
    L.ogs_oracle_preprocess_bwd(
        P, int(degree), M, _p(means3D), _p(fwd["radii"]), _p(shs), _p(fwd["clamped"]), _p(scales), _p(rotations),
        ctypes.c_float(scale_modifier), _p(_c(fwd["cov3D"])), _p(viewmatrix), W, H, _p(campos),
        _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
        _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]), _p(g["dpx_dt"]), _p(g["dpy_dt"]))
    return g


def pinhole_mark_visible(means3D, viewmatrix):
    """checkFrustum (rasterizer_impl.cu:64-77)."""
    means3D = _c(means3D); P = means3D.shape[0]
    present = np.zeros(P, np.uint8)
    lib().ogs_oracle_pinhole_mark_visible(P, _p(means3D), _p(_c(viewmatrix).reshape(-1)), _p(present))
    return present.astype(bool)
