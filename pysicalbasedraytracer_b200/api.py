"""Host-side mirror of the reference's integrator interface over the C-ABI (include/pbrgpu.h).

The reference's boundary for this path is the C++ virtual  Integrator::Render(const Scene&, double&)
(Integrator/Integrator.h:14) with the construction pattern of Main/main.cpp:384-413:

    camera  = CreatePerspectiveCamera(W, H, Inverse(LookAt(eye, look, up)))      Camera/Perspective.cpp:84-104
    sampler = HaltonSampler(spp, Bounds2i((0,0),(W,H)))                          Sampler/Halton.cpp:30-58
    integ   = PathIntegrator(maxDepth, camera, sampler, pixelBounds, rrThreshold, "uniform", &framebuffer)
    integ.Render(scene, timeConsume)                                             Integrator/Integrator.cpp:280-356

The classes below keep those names and argument meanings; `Scene` wraps the flattened triangle soup
(scenes.SceneData) instead of the reference's pointer graph.  Everything that computes runs in
libpbrgpu.so on the GPU: there is no CPU fallback here, and a missing extension raises at import of
the library handle.  (The C++ drop-in for the reference's own build is host/GpuIntegrator.h.)
"""
import ctypes as C
import os
import time

import numpy as np

from . import _abi as A
from . import transform as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PBRGPU_LIB", os.path.join(_HERE, "libpbrgpu.so"))  # override: kernel-variant experiments
_lib = None


class PbrGpuError(RuntimeError):
    pass


def lib():
    """Load libpbrgpu.so (built in-tree by __graft_entry__.build()).  Fails loudly when missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PbrGpuError("libpbrgpu.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.pbrgpu_last_error.restype = C.c_char_p
        L.pbrgpu_init.argtypes = [C.c_int, C.POINTER(vp)]
        L.pbrgpu_shutdown.argtypes = [vp]
        L.pbrgpu_scene_create.argtypes = [vp, C.POINTER(A.SceneDesc), C.POINTER(vp)]
        L.pbrgpu_scene_destroy.argtypes = [vp]
        L.pbrgpu_camera_set.argtypes = [vp, C.POINTER(A.CameraDesc)]
        L.pbrgpu_sampler_set.argtypes = [vp, C.POINTER(A.SamplerDesc)]
        L.pbrgpu_render.argtypes = [vp, C.POINTER(A.RenderDesc), vp, vp, C.POINTER(A.Stats)]
        L.pbrgpu_render_device.argtypes = [vp, C.POINTER(A.RenderDesc), vp, vp, C.POINTER(A.Stats)]
        L.pbrgpu_resolve_film.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, vp, vp]
        L.pbrgpu_debug_dump.argtypes = [vp, C.POINTER(A.DumpDesc), vp, C.c_size_t]
        L.pbrgpu_trace_rays.argtypes = [vp, vp, C.c_int64, C.c_int, vp]
        L.pbrgpu_halton_params.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.pbrgpu_compute_perms.restype = C.c_int64
        L.pbrgpu_compute_perms.argtypes = [vp, C.c_int64]
        L.pbrgpu_scene_info.argtypes = [vp, C.POINTER(C.c_int64)]
        L.pbrgpu_load_3d.argtypes = [C.c_char_p, C.c_float, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int32),
                                     C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]
        L.pbrgpu_texture_eval.argtypes = [vp, C.c_int, vp, C.c_int64, vp]
        L.pbrgpu_load_hdr.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.pbrgpu_free.argtypes = [vp]
        L.pbrgpu_light_eval.argtypes = [vp, C.c_int, vp, C.c_int64, vp]
        L.pbrgpu_resample_pow2.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int64, C.POINTER(C.c_int32)]
        L.pbrgpu_init_multi.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
        L.pbrgpu_comm_unique_id.argtypes = [vp]
        L.pbrgpu_comm_attach.argtypes = [vp, C.c_int, C.c_int, vp]
        L.pbrgpu_ctx_info.argtypes = [vp, C.POINTER(C.c_int32)]
        L.pbrgpu_shard_samples.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.pbrgpu_measure_read_bandwidth.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PbrGpuError(lib().pbrgpu_last_error().decode("utf-8", "replace"))


def device_count():
    return lib().pbrgpu_device_count()


def halton_params(bounds):
    out = (C.c_int32 * 7)()
    lib().pbrgpu_halton_params((C.c_int32 * 4)(*bounds), out)
    return list(out)


def compute_perms():
    n = lib().pbrgpu_compute_perms(None, 0)
    out = np.zeros(n, dtype=np.uint16)
    lib().pbrgpu_compute_perms(out.ctypes.data, n)
    return out


def load_3d(path, scale=20.0):
    """The reference's ".3d" mesh text format (Shape/plyRead.h:17-47; vertices scaled by 20 like plyInfo).
    -> (verts float32 [N,3], tris int32 [M,3])"""
    v, i = C.POINTER(C.c_float)(), C.POINTER(C.c_int32)()
    nv, nt = C.c_int32(0), C.c_int32(0)
    rc = lib().pbrgpu_load_3d(str(path).encode(), scale, C.byref(v), C.byref(nv), C.byref(i), C.byref(nt))
    if rc != 0:
        raise PbrGpuError(f"pbrgpu_load_3d({path}) failed with code {rc}")
    try:
        verts = np.ctypeslib.as_array(v, shape=(nv.value, 3)).copy()
        tris = np.ctypeslib.as_array(i, shape=(nt.value, 3)).copy()
    finally:
        lib().pbrgpu_free(v)
        lib().pbrgpu_free(i)
    return verts, tris


def load_hdr(path, flip_vertically=False):
    """Radiance .hdr environment map as stbi_loadf returns it (Light/InfiniteAreaLight.cpp:17-18; SkyBoxLight.cpp:19-22 loads
    with flip_vertically=True) -> float32 [H,W,3]."""
    p = C.POINTER(C.c_float)()
    w, h = C.c_int32(0), C.c_int32(0)
    rc = lib().pbrgpu_load_hdr(str(path).encode(), int(flip_vertically), C.byref(p), C.byref(w), C.byref(h))
    if rc != 0:
        raise PbrGpuError(f"pbrgpu_load_hdr({path}) failed with code {rc}")
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    finally:
        lib().pbrgpu_free(p)


def comm_unique_id():
    """128-byte NCCL id (rank 0 creates it, the launcher broadcasts it, every rank passes it to Context.attach)."""
    buf = (C.c_ubyte * 128)()
    _check(lib().pbrgpu_comm_unique_id(buf))
    return bytes(buf)


def resample_pow2(img, wrap=A.WRAP_REPEAT):
    """MIPMap's level 0 for an RGB image [H,W,3] whose sides are not powers of two (host-only helper; Lanczos, MIPMap.h:93-144)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    size = (C.c_int32 * 2)()
    _check(lib().pbrgpu_resample_pow2(img.ctypes.data, img.shape[1], img.shape[0], wrap, None, 0, size))
    out = np.zeros((size[1], size[0], 3), dtype=np.float32)
    _check(lib().pbrgpu_resample_pow2(img.ctypes.data, img.shape[1], img.shape[0], wrap, out.ctypes.data, out.size, size))
    return out


def shard_samples(s0, s1, rank, world):
    """The sample sub-range of [s0, s1) that rank `rank` of `world` renders on a multi-GPU context (host-only helper)."""
    out = (C.c_int64 * 2)()
    _check(lib().pbrgpu_shard_samples(s0, s1, rank, world, out))
    return int(out[0]), int(out[1])


class Context:
    """One CUDA device + stream (pbrgpu_ctx), or - `devices` given - a GROUP of devices driven by this process
    (pbrgpu_init_multi): scenes are replicated, Scene.render shards the sample range and reduces the films with NCCL."""

    def __init__(self, device=0, devices=None):
        self.h = C.c_void_p()
        if devices is not None:
            ids = (C.c_int * len(devices))(*devices) if len(devices) else None
            _check(lib().pbrgpu_init_multi(len(devices), ids, C.byref(self.h)))
            device = devices[0] if len(devices) else 0
        else:
            _check(lib().pbrgpu_init(device, C.byref(self.h)))
        self.device = device

    def attach(self, nranks, rank, unique_id):
        """One process per GPU: join rank `rank` of `nranks` (pbrgpu_comm_attach).  Afterwards every rank calls Scene.render /
        render_device with the same arguments; the reduced film arrives on rank 0."""
        assert len(unique_id) == 128
        _check(lib().pbrgpu_comm_attach(self.h, nranks, rank, C.c_char_p(unique_id)))
        return self

    def read_bandwidth(self, nbytes, iters):
        """GB/s of re-reading an nbytes device buffer `iters` times (L2-resident when nbytes is well below the L2 size)."""
        out = C.c_double(0)
        _check(lib().pbrgpu_measure_read_bandwidth(self.h, nbytes, iters, C.byref(out)))
        return out.value

    def info(self):
        out = (C.c_int32 * 4)()
        _check(lib().pbrgpu_ctx_info(self.h, out))
        return dict(n_devices=out[0], rank=out[1], nranks=out[2], nccl_version=out[3])

    def close(self):
        if self.h:
            lib().pbrgpu_shutdown(self.h)
            self.h = C.c_void_p()


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---------------------------------------------------------------------------------------------
# Camera / Sampler / FrameBuffer: POD mirrors of the reference objects the integrator ctor takes
# ---------------------------------------------------------------------------------------------
class Camera:
    def __init__(self, cam_type, width, height, camera_to_world, lens_radius=0.0, focal_distance=0.0, medium=None):
        self.type, self.width, self.height = cam_type, width, height
        # Camera::medium (Camera.h:29).  The reference never delivers it to the integrator: PerspectiveCamera does not
        # set ray->medium (Perspective.cpp:44-80) and OrthographicCamera sets it before `*ray = CameraToWorld(*ray)`,
        # which rebuilds the ray without it (Transform.h:169).  The mirror therefore records the argument and keeps
        # `medium = None`; assign `camera.medium = id` explicitly to make the GPU start its rays inside that medium.
        self.ctor_medium = medium
        self.medium = None
        self.camera_to_world = camera_to_world
        ortho = cam_type == A.CAMERA_ORTHOGRAPHIC
        window = T.screen_window(width, height, ortho)
        c2s = T.orthographic(0, 10) if ortho else T.perspective(90.0, 1e-2, 1000.0)
        self.raster_to_camera = T.raster_to_camera(width, height, c2s, window)
        self.lens_radius, self.focal_distance = lens_radius, focal_distance

    def desc(self):
        d = A.CameraDesc(type=self.type, width=self.width, height=self.height, lens_radius=self.lens_radius,
                         focal_distance=self.focal_distance)
        d.raster_to_camera = (C.c_float * 16)(*[float(x) for x in self.raster_to_camera.m.reshape(-1)])
        d.camera_to_world = (C.c_float * 16)(*[float(x) for x in self.camera_to_world.m.reshape(-1)])
        d.medium_plus1 = 0 if self.medium is None else int(self.medium) + 1
        return d


def CreatePerspectiveCamera(width, height, cam2world, media=None):
    """Camera/Perspective.cpp:84-104 (fov 90, lens radius 0).  `media` is accepted and - as in the reference - has no effect."""
    return Camera(A.CAMERA_PERSPECTIVE, width, height, cam2world, medium=media)


def CreateOrthographicCamera(width, height, cam2world, media=None):
    """Camera/Orthographic.cpp:65-98 (screen window scaled by 2).  `media` is accepted and - as in the reference - has no
    effect (see Camera.__init__)."""
    return Camera(A.CAMERA_ORTHOGRAPHIC, width, height, cam2world, medium=media)


def LookAt(eye, look, up):
    return T.look_at(eye, look, up)


def Inverse(t):
    return t.inverse()


class HaltonSampler:
    """Sampler/Halton.cpp:30-58.  The radical-inverse permutations are generated inside the library
    with the reference's default-seeded PCG32 (RNG.h:93-108)."""

    def __init__(self, samplesPerPixel, sampleBounds, sampleAtPixelCenter=False):
        self.samplesPerPixel = int(samplesPerPixel)
        self.sampleBounds = tuple(int(v) for v in sampleBounds)  # x0, y0, x1, y1
        self.sampleAtPixelCenter = bool(sampleAtPixelCenter)

    def desc(self):
        d = A.SamplerDesc(type=A.SAMPLER_HALTON, sample_at_pixel_center=int(self.sampleAtPixelCenter),
                          samples_per_pixel=self.samplesPerPixel)
        d.sample_bounds = (C.c_int32 * 4)(*self.sampleBounds)
        return d


class SobolSampler:
    """Sobol' (0,2)-sequence sampler (CPU twin: host/SobolSampler.h).  The reference ships the generator tables
    (Sampler/SobolMatrices.h:42-47) but no sampler that uses them, so parity is unpinned; the tables are the
    reference's data and must be supplied by the caller (`matrices32` uint32[n*52], `vdc`/`vdc_inv` uint64[52]
    = VdCSobolMatrices(Inv)[log2(resolution) - 1])."""

    def __init__(self, samplesPerPixel, sampleBounds, matrices32, vdc, vdc_inv):
        spp = 1
        while spp < int(samplesPerPixel):
            spp *= 2
        self.samplesPerPixel = spp  # GlobalSampler(RoundUpPow2(samplesPerPixel))
        self.sampleBounds = tuple(int(v) for v in sampleBounds)
        self.m32 = np.ascontiguousarray(matrices32, dtype=np.uint32)
        self.vdc = np.ascontiguousarray(vdc, dtype=np.uint64)
        self.vdc_inv = np.ascontiguousarray(vdc_inv, dtype=np.uint64)
        assert self.m32.size % 52 == 0 and self.vdc.size == 52 and self.vdc_inv.size == 52

    def desc(self):
        d = A.SamplerDesc(type=A.SAMPLER_SOBOL, samples_per_pixel=self.samplesPerPixel)
        d.sample_bounds = (C.c_int32 * 4)(*self.sampleBounds)
        d.sobol_matrices32 = self.m32.ctypes.data_as(C.POINTER(C.c_uint32))
        d.sobol_n_dims = self.m32.size // 52
        d.sobol_vdc = self.vdc.ctypes.data_as(C.POINTER(C.c_uint64))
        d.sobol_vdc_inv = self.vdc_inv.ctypes.data_as(C.POINTER(C.c_uint64))
        return d


class FrameBuffer:
    """Core/FrameBuffer.h: RGBA8 `ubuffer` (row 0 = bottom, Integrator.cpp:339-341) + float `fbuffer`.
    `fbuffer` here receives the linear per-pixel mean radiance (raster order, y down)."""

    def __init__(self):
        self.width = self.height = self.channals = 0
        self.ubuffer = None
        self.fbuffer = None

    def InitBuffer(self, width=800, height=600, channals=4):
        self.width, self.height, self.channals = width, height, channals
        self.ubuffer = np.zeros((height, width, channals), dtype=np.uint8)
        self.fbuffer = np.zeros((height, width, channals), dtype=np.float32)

    def getUCbuffer(self):
        return self.ubuffer


# ---------------------------------------------------------------------------------------------
class Scene:
    """Flattened Scene (Core/Scene.h) resident in HBM.  `data` is a scenes.SceneData."""

    def __init__(self, data, ctx=None):
        self.ctx = ctx or default_context()
        self.data = data
        d, self._keep = data.desc()
        self.h = C.c_void_p()
        t0 = time.perf_counter()
        _check(lib().pbrgpu_scene_create(self.ctx.h, C.byref(d), C.byref(self.h)))
        self.build_seconds = time.perf_counter() - t0
        self._camera = self._sampler = None

    def info(self):
        out = (C.c_int64 * 8)()
        _check(lib().pbrgpu_scene_info(self.h, out))
        return dict(n_nodes=out[0], n_leaves=out[1], max_leaf=out[2], bytes_nodes=out[3], bytes_tris=out[4],
                    max_depth=out[5], bvh_hash=out[6])

    def set_camera(self, camera):
        d = camera.desc() if hasattr(camera, "desc") else camera
        _check(lib().pbrgpu_camera_set(self.h, C.byref(d)))
        self._camera = camera

    def set_sampler(self, sampler):
        d = sampler.desc() if hasattr(sampler, "desc") else sampler
        _check(lib().pbrgpu_sampler_set(self.h, C.byref(d)))
        self._sampler = sampler

    # ---- parity hooks ----
    def _dump(self, what, region, s0, s1, n_dims, shape, dtype):
        x0, y0, x1, y1 = region
        out = np.zeros(shape, dtype=dtype)
        dd = A.DumpDesc(what=what, x0=x0, y0=y0, x1=x1, y1=y1, s0=s0, s1=s1, n_dims=n_dims)
        _check(lib().pbrgpu_debug_dump(self.h, C.byref(dd), out.ctypes.data, out.nbytes))
        return out

    def dump_sampler(self, region, s0, s1, n_dims):
        x0, y0, x1, y1 = region
        return self._dump(A.DUMP_SAMPLER, region, s0, s1, n_dims, ((y1 - y0) * (x1 - x0), s1 - s0, n_dims), np.float32)

    def dump_primary_rays(self, region, s0, s1):
        x0, y0, x1, y1 = region
        return self._dump(A.DUMP_PRIMARY_RAYS, region, s0, s1, 0, ((y1 - y0) * (x1 - x0), s1 - s0, 8), np.float32)

    def dump_primary_hits(self, region, s0, s1):
        x0, y0, x1, y1 = region
        return self._dump(A.DUMP_PRIMARY_HITS, region, s0, s1, 0, ((y1 - y0) * (x1 - x0), s1 - s0, 4), np.int32)

    def dump_wave_hits(self, region, s0, s1):
        """Primary hits as the PRODUCTION kernels of a render find them (k_raygen -> k_trace): int32 [pix, S, 4]."""
        x0, y0, x1, y1 = region
        return self._dump(A.DUMP_WAVE_HITS, region, s0, s1, 0, ((y1 - y0) * (x1 - x0), s1 - s0, 4), np.int32)

    def dump_hit_differentials(self, region, s0, s1):
        """float32 [H,W,S,8]: hit flag, uv, dudx, dvdx, dudy, dvdy of each camera ray's surface hit."""
        x0, y0, x1, y1 = region
        return self._dump(A.DUMP_HIT_DIFFERENTIALS, region, s0, s1, 0, (y1 - y0, x1 - x0, s1 - s0, 8), np.float32)

    def dump_perms(self, n):
        out = np.zeros(n, dtype=np.uint16)
        dd = A.DumpDesc(what=A.DUMP_PERMS)
        _check(lib().pbrgpu_debug_dump(self.h, C.byref(dd), out.ctypes.data, out.nbytes))
        return out

    def trace(self, rays, any_hit=False, via_wave_kernel=False):
        """Closest / any-hit queries for caller-supplied rays.  via_wave_kernel: through k_trace's own queues (the kernel
        a render launches) instead of the one-thread-per-ray helper loop."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        out = np.zeros((n,) if any_hit else (n, 4), dtype=np.int32)
        mode = int(any_hit) | (A.TRACE_VIA_WAVE_KERNEL if via_wave_kernel else 0)
        _check(lib().pbrgpu_trace_rays(self.h, rays.ctypes.data, n, mode, out.ctypes.data))
        return out

    def texture_eval(self, texture, rows):
        """ImageTexture::Evaluate parity hook: rows [n,8] = u, v, dudx, dvdx, dudy, dvdy, -, - -> rgb [n,3]."""
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 8)
        out = np.zeros((rows.shape[0], 4), dtype=np.float32)
        _check(lib().pbrgpu_texture_eval(self.h, texture, rows.ctypes.data, rows.shape[0], out.ctypes.data))
        return out[:, :3]

    def light_eval(self, light, rows):
        """Light::Sample_Li / Pdf_Li parity hook: rows [n,16] = ref.p, ref.pError, ref.n, u0, u1, query wi, -, - ->
        [n,16] = Li.rgb, pdf, wi, p1, n1, Pdf_Li(ref, query wi), -, -."""
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 16)
        out = np.zeros_like(rows)
        _check(lib().pbrgpu_light_eval(self.h, light, rows.ctypes.data, rows.shape[0], out.ctypes.data))
        return out

    # ---- render ----
    def render(self, integrator=A.INTEGRATOR_PATH, max_depth=5, rr_threshold=1.0,
               light_strategy=A.LIGHTSTRATEGY_UNIFORM, region=None, s0=0, s1=None, flags=0, samples_per_wave=0,
               want_rgba8=False, out=None):
        """-> (rgb_sum float32 [H,W,3], rgba8 or None, Stats).  Host buffers in, host buffers out.
        `out`: optional preallocated float32 [H,W,3] array to receive the sums (like the app's FrameBuffer, which
        is allocated once and reused across Render calls)."""
        if region is None:
            region = (0, 0, self._camera.width, self._camera.height)
        if s1 is None:
            s1 = self._sampler.samplesPerPixel
        x0, y0, x1, y1 = region
        rd = A.RenderDesc(integrator=integrator, max_depth=max_depth, rr_threshold=rr_threshold,
                          light_strategy=light_strategy, x0=x0, y0=y0, x1=x1, y1=y1, s0=s0, s1=s1,
                          samples_per_wave=samples_per_wave, flags=flags)
        if out is None:
            out = np.zeros((y1 - y0, x1 - x0, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == (y1 - y0, x1 - x0, 3) and out.flags["C_CONTIGUOUS"]
        rgba = np.zeros((y1 - y0, x1 - x0, 4), dtype=np.uint8) if want_rgba8 else None
        st = A.Stats()
        _check(lib().pbrgpu_render(self.h, C.byref(rd), out.ctypes.data, rgba.ctypes.data if want_rgba8 else None,
                                   C.byref(st)))
        return out, rgba, st

    def render_device(self, d_rgba_sum_ptr, stream=0, integrator=A.INTEGRATOR_PATH, max_depth=5, rr_threshold=1.0,
                      light_strategy=A.LIGHTSTRATEGY_UNIFORM, region=None, s0=0, s1=None, flags=0, want_stats=True):
        """Render into a DEVICE RGBA-sum accumulator (e.g. a torch tensor's data_ptr())."""
        if region is None:
            region = (0, 0, self._camera.width, self._camera.height)
        if s1 is None:
            s1 = self._sampler.samplesPerPixel
        x0, y0, x1, y1 = region
        rd = A.RenderDesc(integrator=integrator, max_depth=max_depth, rr_threshold=rr_threshold,
                          light_strategy=light_strategy, x0=x0, y0=y0, x1=x1, y1=y1, s0=s0, s1=s1, flags=flags)
        st = A.Stats()
        _check(lib().pbrgpu_render_device(self.h, C.byref(rd), C.c_void_p(d_rgba_sum_ptr), C.c_void_p(stream),
                                          C.byref(st) if want_stats else None))
        return st

    def close(self):
        if self.h:
            lib().pbrgpu_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_STRATEGY = {"uniform": A.LIGHTSTRATEGY_UNIFORM, "power": A.LIGHTSTRATEGY_POWER, "spatial": A.LIGHTSTRATEGY_UNIFORM}


class _GpuSamplerIntegrator:
    """SamplerIntegrator::Render (Integrator.cpp:280-356) on the GPU: pixel x spp loop, per-pixel mean,
    XYZ/sRGB/gamma, 8-bit with Y flip.  Deviation: iterates x in [0,W), y in [0,H)."""
    kind = A.INTEGRATOR_PATH

    def __init__(self, camera, sampler, pixelBounds, framebuffer=None):
        self.camera, self.sampler, self.pixelBounds, self.framebuffer = camera, sampler, tuple(pixelBounds), framebuffer
        self.IntegratorRenderTime = 0.0
        self.stats = None

    def _render_args(self):
        return {}

    def Render(self, scene):
        """Returns timeConsume (seconds), like the reference's out-parameter."""
        t0 = time.perf_counter()
        scene.set_camera(self.camera)
        scene.set_sampler(self.sampler)
        rgb_sum, rgba8, st = scene.render(integrator=self.kind, region=self.pixelBounds, want_rgba8=True,
                                          **self._render_args())
        self.stats = st
        self.rgb_mean = rgb_sum / np.float32(self.sampler.samplesPerPixel)
        fb = self.framebuffer
        if fb is not None and fb.ubuffer is not None:
            x0, y0, x1, y1 = self.pixelBounds
            fb.ubuffer[fb.height - y1:fb.height - y0, x0:x1, :] = rgba8
            fb.fbuffer[y0:y1, x0:x1, :3] = self.rgb_mean
        self.IntegratorRenderTime = time.perf_counter() - t0
        return self.IntegratorRenderTime


class PathIntegrator(_GpuSamplerIntegrator):
    """Integrator/PathIntegrator.h:14-18 - same constructor arguments."""
    kind = A.INTEGRATOR_PATH

    def __init__(self, maxDepth, camera, sampler, pixelBounds, rrThreshold=1.0, lightSampleStrategy="spatial",
                 framebuffer=None):
        super().__init__(camera, sampler, pixelBounds, framebuffer)
        self.maxDepth, self.rrThreshold, self.lightSampleStrategy = maxDepth, rrThreshold, lightSampleStrategy

    def _render_args(self):
        return dict(max_depth=self.maxDepth, rr_threshold=self.rrThreshold,
                    light_strategy=_STRATEGY.get(self.lightSampleStrategy, A.LIGHTSTRATEGY_UNIFORM))


class VolPathIntegrator(PathIntegrator):
    """Integrator/VolPathIntegrator.h:11-19 - same constructor arguments as PathIntegrator; handles participating
    media (SceneData.add_medium / medium_interface) and surfaces without a material."""
    kind = A.INTEGRATOR_VOLPATH


class WhittedIntegrator(_GpuSamplerIntegrator):
    """Integrator/WhittedIntegrator.h:8-11."""
    kind = A.INTEGRATOR_WHITTED

    def __init__(self, maxDepth, camera, sampler, pixelBounds, framebuffer=None):
        super().__init__(camera, sampler, pixelBounds, framebuffer)
        self.maxDepth = maxDepth

    def _render_args(self):
        return dict(max_depth=self.maxDepth)


class DirectLightingIntegrator(_GpuSamplerIntegrator):
    """Integrator/DirectLightingIntegrator.h:10-17.  strategy: "UniformSampleAll" | "UniformSampleOne"."""

    def __init__(self, strategy, maxDepth, camera, sampler, pixelBounds, framebuffer=None):
        super().__init__(camera, sampler, pixelBounds, framebuffer)
        self.maxDepth = maxDepth
        self.kind = A.INTEGRATOR_DIRECT_ALL if strategy == "UniformSampleAll" else A.INTEGRATOR_DIRECT_ONE

    def _render_args(self):
        return dict(max_depth=self.maxDepth)
