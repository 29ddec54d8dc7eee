"""ORACLE (test infrastructure) -- numpy front end of oracle/geometry_ops.c.

Mirrors the Python signatures of the reference's op binding
(lmbspecialops/python/lmbspecialops/__init__.py:45-58,296-308; documented in
lmbspecialops/doc/lmbspecialops_doc.md) on numpy arrays, float32 or float64.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import ctypes
import os
import subprocess
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_ops.so")


def build(force=False):
    """Compile the C restatement with gcc (a few hundred ms)."""
    src = [os.path.join(_HERE, f) for f in ("geometry_ops.c", "geometry_ops_impl.h")]
    if (not force and os.path.isfile(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


_ROT = {"matrix": 0, "quaternion": 1, "angleaxis3": 2}


def _sfx(dtype):
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError("oracle ops take float32 or float64, got %s" % dtype)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def warp2d(input, displacements, normalized=False, border_mode="clamp", border_value=0.0):
    inp = np.asarray(input)
    sfx, cty = _sfx(inp.dtype)
    inp = _c(inp, inp.dtype)
    disp = _c(displacements, inp.dtype)
    if inp.ndim < 2 or disp.ndim < 3 or disp.shape[-3] != 2 or disp.shape[-2:] != inp.shape[-2:]:
        raise ValueError("warp2d: bad shapes %s %s" % (inp.shape, disp.shape))
    x, y = inp.shape[-1], inp.shape[-2]
    z = inp.shape[-3] if inp.ndim >= 3 else 1
    w = int(np.prod(inp.shape[:-3])) if inp.ndim > 3 else 1
    out = np.empty_like(inp)
    getattr(lib(), "oracle_warp2d" + sfx)(
        _p(out), _p(inp), _p(disp), x, y, z, w, int(bool(normalized)),
        1 if border_mode == "clamp" else 2, cty(border_value))
    return out


def rotation_matrix(rotation, rotation_format="angleaxis3"):
    rot = np.asarray(rotation)
    sfx, _ = _sfx(rot.dtype)
    step = {0: 9, 1: 4, 2: 3}[_ROT[rotation_format]]
    rot = _c(rot, rot.dtype).reshape(-1, step)
    out = np.empty((rot.shape[0], 3, 3), rot.dtype)
    getattr(lib(), "oracle_rotation_matrix" + sfx)(_p(out), _p(rot), _ROT[rotation_format], rot.shape[0])
    return out


def _pose_args(n, dtype, intrinsics, rotation, translation, rotation_format):
    step = {0: 9, 1: 4, 2: 3}[_ROT[rotation_format]]
    k = _c(intrinsics, dtype).reshape(-1, 4)
    r = _c(rotation, dtype).reshape(-1, step)
    t = _c(translation, dtype).reshape(-1, 3)
    if not (k.shape[0] == r.shape[0] == t.shape[0] == n):
        raise ValueError("Dimensions must be equal")
    return k, r, t


def depth_to_flow(depth, intrinsics, rotation, translation, rotation_format="angleaxis3",
                  inverse_depth=False, normalize_flow=False):
    d = np.asarray(depth)
    sfx, _ = _sfx(d.dtype)
    d = _c(d, d.dtype)
    y, x = d.shape[-2:]
    n = int(np.prod(d.shape[:-2])) if d.ndim > 2 else 1
    k, r, t = _pose_args(n, d.dtype, intrinsics, rotation, translation, rotation_format)
    out = np.empty((n, 2, y, x), d.dtype)
    getattr(lib(), "oracle_depth_to_flow" + sfx)(
        _p(out), _p(d), _p(k), _p(r), _p(t), x, y, n, _ROT[rotation_format],
        int(bool(inverse_depth)), int(bool(normalize_flow)))
    return out


def flow_to_depth2(flow, intrinsics, rotation, translation, rotation_format="angleaxis3",
                   inverse_depth=False, normalized_flow=False):
    f = np.asarray(flow)
    sfx, _ = _sfx(f.dtype)
    f = _c(f, f.dtype)
    if f.ndim < 3 or f.shape[-3] != 2:
        raise ValueError("flow must be [..,2,H,W]")
    y, x = f.shape[-2:]
    n = int(np.prod(f.shape[:-3])) if f.ndim > 3 else 1
    k, r, t = _pose_args(n, f.dtype, intrinsics, rotation, translation, rotation_format)
    out = np.empty((n, 1, y, x), f.dtype)
    getattr(lib(), "oracle_flow_to_depth" + sfx)(
        _p(out), _p(f), _p(k), _p(r), _p(t), x, y, n, _ROT[rotation_format],
        int(bool(inverse_depth)), int(bool(normalized_flow)))
    return out


def flow_to_depth(flow, intrinsics, rotation, translation, rotation_format=None, inverse_depth=None,
                  normalized_flow=None, name=None, nowarning=False):
    """Deprecated twin (lmbspecialops/__init__.py:296-308); numerically == flow_to_depth2."""
    if not nowarning:
        warnings.warn("flow_to_depth has incorrect behaviour but is kept for compatibility. "
                      "Please use flow_to_depth2", DeprecationWarning, stacklevel=2)
    return flow_to_depth2(flow, intrinsics, rotation, translation,
                          rotation_format or "angleaxis3", bool(inverse_depth), bool(normalized_flow))


def leaky_relu(input, leak=0.1):
    a = np.asarray(input)
    sfx, cty = _sfx(a.dtype)
    a = _c(a, a.dtype)
    out = np.empty_like(a)
    getattr(lib(), "oracle_leaky_relu" + sfx)(_p(out), _p(a), ctypes.c_long(a.size), cty(np.float32(leak)))
    return out


def median3x3_downsample(input):
    a = np.asarray(input)
    sfx, _ = _sfx(a.dtype)
    a = _c(a, a.dtype)
    if a.ndim < 2:
        raise ValueError("rank must be at least 2")
    y, x = a.shape[-2:]
    z = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    out = np.empty(a.shape[:-2] + ((y + 1) // 2, (x + 1) // 2), a.dtype)
    getattr(lib(), "oracle_median3x3_downsample" + sfx)(_p(out), _p(a), ctypes.c_long(z), y, x)
    return out


def scale_invariant_gradient(input, deltas=(1,), weights=(1.0,), epsilon=0.001):
    a = np.asarray(input)
    sfx, cty = _sfx(a.dtype)
    a = _c(a, a.dtype)
    if len(deltas) != len(weights):
        raise ValueError("The size of the deltas and weights vectors must be the same")
    y, x = a.shape[-2:]
    z = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    out = np.empty((z, 2, y, x), a.dtype)
    dl = np.asarray(deltas, np.int32)
    # weights and epsilon are float attrs in the reference, converted to T (scaleinvariantgradient.cc:109-113)
    wt = np.asarray(weights, np.float32).astype(a.dtype)
    getattr(lib(), "oracle_scale_invariant_gradient" + sfx)(
        _p(out), _p(a), x, y, ctypes.c_long(z), _p(dl), _p(wt), len(dl), cty(np.float32(epsilon)))
    return out


def depth_to_normals(depth, intrinsics, inverse_depth=False):
    """depthtonormals.cc:147-238: normal map [N,3,H,W] of a depth map (leading dims collapse into N)."""
    a = np.asarray(depth)
    sfx, _ = _sfx(a.dtype)
    a = _c(a, a.dtype)
    if a.ndim < 2:
        raise ValueError("rank must be at least 2")
    y, x = a.shape[-2:]
    z = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    k = np.ascontiguousarray(np.broadcast_to(np.asarray(intrinsics, a.dtype).reshape(-1, 4), (z, 4)) if np.asarray(intrinsics).size == 4
                             else np.asarray(intrinsics, a.dtype).reshape(z, 4))
    out = np.empty((z, 3, y, x), a.dtype)
    getattr(lib(), "oracle_depth_to_normals" + sfx)(_p(out), _p(a), _p(k), x, y, ctypes.c_long(z), int(bool(inverse_depth)))
    return out
