"""TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/_ref/libpbr_ref.so (the compiled, unmodified
reference + oracle/ref_harness.cpp).  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package."""
import ctypes as C
import os

import numpy as np

from pysicalbasedraytracer_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpbr_ref.so")


class RefRenderStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("seconds", C.c_double), ("threads", C.c_int)]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.POINTER(A.SceneDesc)]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_scene_bvh_seconds.restype = C.c_double
        L.ref_scene_bvh_seconds.argtypes = [C.c_void_p]
        L.ref_scene_bound.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.ref_camera_create.restype = C.c_void_p
        L.ref_camera_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.POINTER(C.c_float)]
        L.ref_camera_destroy.argtypes = [C.c_void_p]
        L.ref_camera_create_lens.restype = C.c_void_p
        L.ref_camera_create_lens.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             C.POINTER(C.c_float), C.c_float, C.c_float]
        L.ref_camera_desc.argtypes = [C.c_void_p, C.POINTER(A.CameraDesc)]
        L.ref_camera_hit_differentials.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        L.ref_texture_kat.argtypes = [C.POINTER(A.Texture), C.c_void_p, C.c_int, C.c_void_p]
        L.ref_load_hdr.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.ref_medium_kat.argtypes = [C.POINTER(A.Medium), C.c_void_p, C.c_int, C.c_void_p]
        L.ref_hg_kat.argtypes = [C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_camera_set_medium.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_sampler_create.restype = C.c_void_p
        L.ref_sampler_create.argtypes = [C.c_int] * 5
        L.ref_sampler_destroy.argtypes = [C.c_void_p]
        L.ref_halton_params.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.ref_perms.restype = C.c_int64
        L.ref_perms.argtypes = [C.c_void_p, C.c_int64]
        L.ref_radical_inverse.restype = C.c_float
        L.ref_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.ref_scrambled_radical_inverse.restype = C.c_float
        L.ref_scrambled_radical_inverse.argtypes = [C.c_int, C.c_uint64]
        L.ref_prime.argtypes = [C.c_int]
        L.ref_prime_sum.argtypes = [C.c_int]
        L.ref_sampler_dump.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.ref_halton_index.restype = C.c_int64
        L.ref_halton_index.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64]
        L.ref_primary_rays.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_int64, C.c_int64, C.c_void_p]
        L.ref_primary_hits.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_int64, C.c_int64, C.c_void_p]
        L.ref_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.ref_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(A.RenderDesc),
                                 C.c_int, C.c_void_p, C.POINTER(RefRenderStats)]
        L.ref_render_shipped.restype = C.c_double
        L.ref_render_shipped.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                         C.c_void_p, C.POINTER(RefRenderStats)]
        L.ref_tonemap.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_ply_read.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.ref_sampler_create_typed.restype = C.c_void_p
        L.ref_sampler_create_typed.argtypes = [C.c_int] * 6
        L.ref_sobol_tables.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_render_typed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(A.RenderDesc),
                                       C.c_int, C.c_void_p, C.POINTER(RefRenderStats)]
        L.ref_render_shipped_integrator.restype = C.c_double
        L.ref_render_shipped_integrator.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                    C.c_void_p, C.POINTER(RefRenderStats)]
        _lib = L
    return _lib


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class RefScene:
    def __init__(self, scene_data):
        d, self._keep = scene_data.desc()
        self.h = lib().ref_scene_create(C.byref(d))

    def bvh_seconds(self):
        return lib().ref_scene_bvh_seconds(self.h)

    def bound(self):
        out = (C.c_float * 6)()
        lib().ref_scene_bound(self.h, out)
        return np.array(out[:], dtype=np.float32)

    def light_eval(self, light, rows):
        """The reference's Light::Sample_Li / Pdf_Li (row layout of pbrgpu_light_eval)."""
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 16)
        out = np.zeros_like(rows)
        lib().ref_light_eval(C.c_void_p(self.h), light, rows.ctypes.data, rows.shape[0], out.ctypes.data)
        return out

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        out = np.zeros((n,) if any_hit else (n, 4), dtype=np.int32)
        lib().ref_trace_rays(self.h, rays.ctypes.data, n, int(any_hit), out.ctypes.data)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_scene_destroy(self.h)
            self.h = None


class RefCamera:
    def __init__(self, cam_type, W, H, eye, look, up, lens_radius=0.0, focal_distance=0.0):
        if lens_radius > 0:
            self.h = lib().ref_camera_create_lens(cam_type, W, H, _f3(eye), _f3(look), _f3(up), lens_radius, focal_distance)
        else:
            self.h = lib().ref_camera_create(cam_type, W, H, _f3(eye), _f3(look), _f3(up))
        self.W, self.H = W, H

    def desc(self):
        d = A.CameraDesc()
        lib().ref_camera_desc(self.h, C.byref(d))
        return d

    def set_medium(self, rscene, medium, honour=False):
        """Camera::medium = medium id of `rscene` (or None).  The reference loses it on the way to the world-space ray
        (Transform.h:169); honour=True re-attaches it in the harness loop (see ref_harness.cpp)."""
        lib().ref_camera_set_medium(self.h, rscene.h, -1 if medium is None else int(medium), int(honour))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_camera_destroy(self.h)
            self.h = None


class RefSampler:
    def __init__(self, spp, bounds, sampler_type=A.SAMPLER_HALTON):
        self.h = lib().ref_sampler_create_typed(sampler_type, spp, *bounds)
        self.spp, self.bounds, self.type = spp, tuple(bounds), sampler_type

    def params(self):
        out = (C.c_int32 * 7)()
        lib().ref_halton_params(self.h, out)
        return list(out)

    def dump(self, region, s0, s1, n_dims):
        x0, y0, x1, y1 = region
        out = np.zeros(((y1 - y0) * (x1 - x0), s1 - s0, n_dims), dtype=np.float32)
        lib().ref_sampler_dump(self.h, x0, y0, x1, y1, s0, s1, n_dims, out.ctypes.data)
        return out

    def index(self, x, y, s):
        return lib().ref_halton_index(self.h, x, y, s)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_sampler_destroy(self.h)
            self.h = None


def perms():
    n = lib().ref_perms(None, 0)
    out = np.zeros(n, dtype=np.uint16)
    lib().ref_perms(out.ctypes.data, n)
    return out


def primary_rays(cam, smp, region, s0, s1):
    x0, y0, x1, y1 = region
    out = np.zeros(((y1 - y0) * (x1 - x0), s1 - s0, 8), dtype=np.float32)
    lib().ref_primary_rays(cam.h, smp.h, x0, y0, x1, y1, s0, s1, out.ctypes.data)
    return out


def primary_hits(scene, cam, smp, region, s0, s1):
    x0, y0, x1, y1 = region
    out = np.zeros(((y1 - y0) * (x1 - x0), s1 - s0, 4), dtype=np.int32)
    lib().ref_primary_hits(scene.h, cam.h, smp.h, x0, y0, x1, y1, s0, s1, out.ctypes.data)
    return out


def sobol_tables(bounds):
    """(matrices32 uint32[1024*52], vdc uint64[52], vdc_inv uint64[52]) of the compiled reference for these sample bounds."""
    res, lg = 1, 0
    while res < max(bounds[2] - bounds[0], bounds[3] - bounds[1]):
        res, lg = res * 2, lg + 1
    m32 = np.zeros(1024 * 52, dtype=np.uint32)
    vdc = np.zeros(52, dtype=np.uint64)
    vdi = np.zeros(52, dtype=np.uint64)
    lib().ref_sobol_tables(m32.ctypes.data, lg, vdc.ctypes.data, vdi.ctypes.data)
    return m32, vdc, vdi


def render(scene, cam, spp, sample_bounds, integrator=A.INTEGRATOR_PATH, max_depth=5, rr_threshold=1.0,
           light_strategy=A.LIGHTSTRATEGY_UNIFORM, region=None, s0=0, s1=None, threads=0, sampler_type=A.SAMPLER_HALTON):
    """-> (rgb_sum float32 [H,W,3], RefRenderStats)"""
    if region is None:
        region = (0, 0, cam.W, cam.H)
    if s1 is None:
        s1 = spp
    x0, y0, x1, y1 = region
    rd = A.RenderDesc(integrator=integrator, max_depth=max_depth, rr_threshold=rr_threshold,
                      light_strategy=light_strategy, x0=x0, y0=y0, x1=x1, y1=y1, s0=s0, s1=s1)
    out = np.zeros((y1 - y0, x1 - x0, 3), dtype=np.float32)
    st = RefRenderStats()
    sb = (C.c_int32 * 4)(*sample_bounds)
    rc = lib().ref_render_typed(scene.h, cam.h, sampler_type, spp, sb, C.byref(rd), threads, out.ctypes.data, C.byref(st))
    if rc != 0:
        raise RuntimeError("ref_render failed")
    return out, st


def render_shipped(scene, cam, spp, W, H, max_depth, rr, integrator=A.INTEGRATOR_PATH):
    """The reference's OWN SamplerIntegrator::Render (4 OpenMP threads, square images) -> FrameBuffer bytes."""
    out = np.zeros((H, W, 4), dtype=np.uint8)
    st = RefRenderStats()
    t = lib().ref_render_shipped_integrator(scene.h, cam.h, integrator, spp, W, H, max_depth, rr, out.ctypes.data, C.byref(st))
    return out, st, t


def tonemap(rgb_mean):
    a = np.ascontiguousarray(rgb_mean, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((a.shape[0], 4), dtype=np.uint8)
    lib().ref_tonemap(a.ctypes.data, a.shape[0], out.ctypes.data)
    return out


_dropin = None


def dropin_lib():
    """oracle/_ref/libdropin_test.so: the C++ drop-in test (the only oracle-side object that links libpbrgpu.so)."""
    global _dropin
    if _dropin is None:
        lib()  # libpbr_ref.so first (libdropin_test.so needs its symbols)
        D = C.CDLL(os.path.join(os.path.dirname(LIB_PATH), "libdropin_test.so"))
        D.ref_gpu_dropin_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                            C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int]
        _dropin = D
    return _dropin


def gpu_dropin_render(scene, cam, spp, W, H, max_depth, rr, strategy="uniform", integrator=A.INTEGRATOR_PATH, n_devices=1):
    """The reference's own object graph rendered through the C++ drop-in GpuPathIntegrator (needs a GPU).
    n_devices > 1: GpuPathIntegrator::setDevices({0..n-1}) - one process driving several GPUs (pbrgpu_init_multi)."""
    rgba = np.zeros((H, W, 4), dtype=np.uint8)
    rgb_sum = np.zeros((H, W, 3), dtype=np.float32)
    sec = C.c_double(0)
    rc = dropin_lib().ref_gpu_dropin_render(scene.h, cam.h, integrator, spp, W, H, max_depth, rr, strategy.encode(), rgba.ctypes.data,
                                            rgb_sum.ctypes.data, C.byref(sec), n_devices)
    if rc != 0:
        raise RuntimeError("GpuPathIntegrator::Render failed (see stderr)")
    return rgba, rgb_sum, sec.value


def ply_read(path, cap_verts=1 << 20, cap_tris=1 << 21):
    """The reference's own plyInfo loader (Shape/plyRead.h) -> (verts [N,3] float32, tris [M,3] int32)."""
    v = np.zeros((cap_verts, 3), dtype=np.float32)
    t = np.zeros((cap_tris, 3), dtype=np.int32)
    nv, nt = C.c_int32(0), C.c_int32(0)
    rc = lib().ref_ply_read(str(path).encode(), v.ctypes.data, cap_verts, t.ctypes.data, cap_tris, C.byref(nv), C.byref(nt))
    if rc != 0:
        raise RuntimeError("ref_ply_read: buffers too small")
    return v[:nv.value].copy(), t[:nt.value].copy()


def medium_kat(sigma_a, sigma_s, g, rows):
    """The reference's HomogeneousMedium::Sample / ::Tr on rays (0, d, tMax) with sampler values (u0, u1); see ref_harness.cpp."""
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 8)
    out = np.zeros_like(rows)
    m = A.Medium(type=0, g=g, sigma_a=_f3(sigma_a), sigma_s=_f3(sigma_s))
    lib().ref_medium_kat(C.byref(m), rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out


def hg_kat(g, rows):
    """The reference's HenyeyGreenstein::Sample_p / ::p; see ref_harness.cpp."""
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 8)
    out = np.zeros_like(rows)
    lib().ref_hg_kat(g, rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out


def load_hdr(path, flip=False, cap=1 << 24):
    """stbi_loadf of the reference on a Radiance .hdr file -> float32 [H,W,C]."""
    out = np.zeros(cap, dtype=np.float32)
    w, h, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    rc = lib().ref_load_hdr(str(path).encode(), int(flip), out.ctypes.data, cap, C.byref(w), C.byref(h), C.byref(c))
    if rc != 0:
        raise RuntimeError(f"ref_load_hdr failed ({rc})")
    return out[: w.value * h.value * c.value].reshape(h.value, w.value, c.value).copy()


def texture_kat(tex, rows):
    """ImageTexture::Evaluate of the reference for rows [n,8] = u, v, dudx, dvdx, dudy, dvdy, -, - -> rgb [n,3].
    tex: (A.Texture, texel array) as stored in SceneData.textures."""
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 8)
    out = np.zeros((rows.shape[0], 4), dtype=np.float32)
    lib().ref_texture_kat(C.byref(tex[0]), rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out[:, :3]


def camera_hit_differentials(scene, cam, spp, px, py, s):
    """-> float32[16]: hit flag, uv, dudx, dvdx, dudy, dvdy, dpdu, dpdv of the camera ray of (pixel, sample)."""
    out = np.zeros(16, dtype=np.float32)
    lib().ref_camera_hit_differentials(scene.h, cam.h, spp, cam.W, cam.H, px, py, s, out.ctypes.data)
    return out
