"""TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/restate.c (the plain-C restatement of the
bit-exact parts of the reference path).  Imported by tests/ only; never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "restate.c")
LIB_PATH = os.path.join(_HERE, "_build", "librestate.so")
_lib = None


def build():
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB_PATH, SRC, "-lm"],
                       check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.rs_perms.restype = C.c_int64
        L.rs_perms.argtypes = [C.c_void_p, C.c_int64]
        L.rs_radical_inverse.restype = C.c_float
        L.rs_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.rs_scrambled_radical_inverse.restype = C.c_float
        L.rs_scrambled_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.rs_halton_params.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.rs_halton_index.restype = C.c_int64
        L.rs_halton_index.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int64]
        L.rs_sampler_dump.argtypes = [C.POINTER(C.c_int32)] + [C.c_int] * 4 + [C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.rs_primary_ray.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int64,
                                     C.c_void_p]
        L.rs_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                               C.POINTER(C.c_int64)]
        L.rs_tonemap.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.rs_medium_kat.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.rs_hg_kat.argtypes = [C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def perms():
    n = lib().rs_perms(None, 0)
    out = np.zeros(n, dtype=np.uint16)
    lib().rs_perms(out.ctypes.data, n)
    return out


def halton_params(bounds):
    out = (C.c_int32 * 7)()
    lib().rs_halton_params((C.c_int32 * 4)(*bounds), out)
    return list(out)


def halton_index(bounds, x, y, s):
    p7 = (C.c_int32 * 7)(*halton_params(bounds))
    return lib().rs_halton_index(p7, x, y, s)


def sampler_dump(bounds, region, s0, s1, n_dims):
    x0, y0, x1, y1 = region
    out = np.zeros(((y1 - y0) * (x1 - x0), s1 - s0, n_dims), dtype=np.float32)
    lib().rs_sampler_dump((C.c_int32 * 4)(*bounds), x0, y0, x1, y1, s0, s1, n_dims, out.ctypes.data)
    return out


def primary_rays(cam_desc, bounds, region, s0, s1):
    x0, y0, x1, y1 = region
    p7 = (C.c_int32 * 7)(*halton_params(bounds))
    r2c = np.array(cam_desc.raster_to_camera[:], dtype=np.float32)
    c2w = np.array(cam_desc.camera_to_world[:], dtype=np.float32)
    out = np.zeros(((y1 - y0) * (x1 - x0), s1 - s0, 8), dtype=np.float32)
    k = 0
    for y in range(y0, y1):
        for x in range(x0, x1):
            for s in range(s0, s1):
                lib().rs_primary_ray(cam_desc.type, r2c.ctypes.data, c2w.ctypes.data, p7, x, y, s,
                                     out[k, s - s0].ctypes.data)
            k += 1
    return out


def trace(verts, indices, rays, any_hit=False):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    n = rays.shape[0]
    out = np.zeros((n,) if any_hit else (n, 4), dtype=np.int32)
    ties = C.c_int64(0)
    lib().rs_trace(verts.ctypes.data, indices.ctypes.data, indices.shape[0], rays.ctypes.data, n, int(any_hit),
                   out.ctypes.data, C.byref(ties))
    return out, ties.value


def tonemap(rgb_mean):
    a = np.ascontiguousarray(rgb_mean, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((a.shape[0], 4), dtype=np.uint8)
    lib().rs_tonemap(a.ctypes.data, a.shape[0], out.ctypes.data)
    return out


def medium_kat(sigma_a, sigma_s, g, rows):
    """rows [n,8] = d.xyz, tMax, u0, u1, -, - -> [n,8] = weight rgb, mi.p on d's dominant axis (or tMax), sampled, Tr rgb"""
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 8)
    out = np.zeros_like(rows)
    f3 = C.c_float * 3
    lib().rs_medium_kat(f3(*sigma_a), f3(*sigma_s), g, rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out


def hg_kat(g, rows):
    """rows [n,8] = wo.xyz, u0, u1, wi2.xyz -> [n,8] = Sample_p wi.xyz, p; p(wo, wi2)"""
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 8)
    out = np.zeros_like(rows)
    lib().rs_hg_kat(g, rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out
