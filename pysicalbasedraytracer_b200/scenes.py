"""Synthetic scenes of the BASELINE.json configs (the reference ships no assets).

Every generator returns a SceneData whose arrays are float32/int32 numpy buffers; the same bytes are
handed to the GPU library (pbrgpu_scene_desc) and - in tests - to the compiled reference, so both
sides see byte-identical inputs.  Geometry constants for the Cornell box and its area light follow
the commented-out scene in the reference's Main/main.cpp:285-324 and :353-375.
"""
import ctypes as C

import numpy as np

from . import _abi as A

F = np.float32


def _f3(v):
    return (C.c_float * 3)(*[float(F(x)) for x in v])


def matte(kd, sigma=0.0, kd_texture=None):
    """kd_texture: id from SceneData.add_texture - Kd becomes that ImageTexture (kd is then ignored)."""
    return A.Material(type=A.MAT_MATTE, kd=_f3(kd), sigma=sigma, kd_texture_plus1=0 if kd_texture is None else kd_texture + 1)


def mirror(kr):
    return A.Material(type=A.MAT_MIRROR, kr=_f3(kr))


def glass(kr, kt, index, u_rough=0.0, v_rough=0.0, remap=False):
    return A.Material(type=A.MAT_GLASS, kr=_f3(kr), kt=_f3(kt), eta=_f3([index, index, index]), u_rough=u_rough,
                      v_rough=v_rough, remap_roughness=int(remap))


def metal(eta, k, u_rough, v_rough, remap=False):
    return A.Material(type=A.MAT_METAL, eta=_f3(eta), k=_f3(k), u_rough=u_rough, v_rough=v_rough,
                      remap_roughness=int(remap))


def plastic(kd, ks, rough, remap=True, kd_texture=None, ks_texture=None, rough_texture=None):
    return A.Material(type=A.MAT_PLASTIC, kd=_f3(kd), kr=_f3(ks), u_rough=rough, v_rough=rough,
                      remap_roughness=int(remap), kd_texture_plus1=0 if kd_texture is None else kd_texture + 1,
                      ks_texture_plus1=0 if ks_texture is None else ks_texture + 1,
                      rough_texture_plus1=0 if rough_texture is None else rough_texture + 1)


class SceneData:
    """Triangle soup + material / light tables (the flattened form of Scene/Primitive/Light graphs)."""

    def __init__(self):
        self._verts = []
        self._normals = []
        self._uvs = []
        self._tangents = []
        self._mesh_flags = []
        self._tris = []
        self._tri_mat = []
        self._tri_light = []
        self._tri_med_in = []
        self._tri_med_out = []
        self.media = []
        self.spheres = []   # analytic spheres (A.Sphere); primitive ids n_tris .. n_tris + len(spheres) - 1
        self.textures = []  # (A.Texture, texel array) pairs
        self.materials = []
        self.lights = []
        self.env = None  # (H, W, 3) float32
        self.tri_flags = None  # optional uint8[n_tris]: bit0 = reverseOrientation ^ transformSwapsHandedness
        self._nverts = 0
        self._ntris = 0
        self._frozen = None

    # ---- construction ----
    def add_material(self, m):
        self.materials.append(m)
        return len(self.materials) - 1

    def add_texture(self, texels, do_trilinear=False, max_anisotropy=8.0, wrap=A.WRAP_REPEAT, su=1.0, sv=1.0, du=0.0, dv=0.0):
        """ImageTexture<RGBSpectrum, Spectrum>(UVMapping2D(su, sv, du, dv), ..., doTrilinear, maxAniso, wrap) over the RGB
        image `texels` [H,W,3] (any size; the values the MIPMap receives, Texture/ImageTexture.cpp:88-95) -> id."""
        tx = np.ascontiguousarray(texels, dtype=F)
        if tx.ndim == 2:  # an ImageTexture<float, float>: the float texels replicated in R, G, B
            tx = np.ascontiguousarray(np.repeat(tx[:, :, None], 3, axis=2))
        assert tx.ndim == 3 and tx.shape[2] == 3
        t = A.Texture(width=tx.shape[1], height=tx.shape[0], texels=tx.ctypes.data_as(C.POINTER(C.c_float)),
                      do_trilinear=int(do_trilinear), max_anisotropy=max_anisotropy, wrap=wrap, su=su, sv=sv, du=du, dv=dv)
        self.textures.append((t, tx))
        return len(self.textures) - 1

    def add_medium(self, sigma_a, sigma_s, g):
        """HomogeneousMedium(sigma_a, sigma_s, g)  (Media/HomogeneousMedium.h:12-16) -> medium id."""
        self.media.append(A.Medium(type=A.MEDIUM_HOMOGENEOUS, g=float(g), sigma_a=_f3(sigma_a), sigma_s=_f3(sigma_s)))
        return len(self.media) - 1

    def add_mesh(self, verts, tris, material, emit=None, two_sided=False, n_samples=1, normals=None, uvs=None,
                 reverse_orientation=False, medium_interface=None, tangents=None):
        """material: int id or per-triangle int array; None / -1 = no material (a pure medium boundary, VolPath only).
        emit: RGB Lemit -> one DiffuseAreaLight per triangle.
        normals / uvs / tangents: optional per-vertex shading normals, uv and tangents (TriangleMesh::n / ::uv / ::s; world space).
        medium_interface: (inside, outside) medium ids or None each  (MediumInterface, Media/Medium.h:51-59)."""
        if material is None:
            material = -1
        verts = np.ascontiguousarray(verts, dtype=F).reshape(-1, 3)
        tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        n = tris.shape[0]
        self._normals.append(np.zeros_like(verts) if normals is None else np.ascontiguousarray(normals, dtype=F).reshape(-1, 3))
        self._uvs.append(np.zeros((verts.shape[0], 2), dtype=F) if uvs is None else np.ascontiguousarray(uvs, dtype=F).reshape(-1, 2))
        self._tangents.append(np.zeros_like(verts) if tangents is None else np.ascontiguousarray(tangents, dtype=F).reshape(-1, 3))
        self._mesh_flags.append(np.full(n, (3 if reverse_orientation else 0) | (4 if normals is not None else 0) |
                                        (8 if uvs is not None else 0) | (16 if tangents is not None else 0), dtype=np.uint8))
        mats = np.full(n, material, dtype=np.int32) if np.isscalar(material) else np.asarray(material, dtype=np.int32)
        lights = np.full(n, -1, dtype=np.int32)
        if emit is not None:
            for i in range(n):
                lights[i] = len(self.lights)
                self.lights.append(A.Light(type=A.LIGHT_AREA_TRI, tri=self._ntris + i, two_sided=int(two_sided),
                                           n_samples=n_samples, L=_f3(emit)))
        self._verts.append(verts)
        self._tris.append(tris + self._nverts)
        self._tri_mat.append(mats)
        self._tri_light.append(lights)
        mi = medium_interface if medium_interface is not None else (None, None)
        self._tri_med_in.append(np.full(n, -1 if mi[0] is None else mi[0], dtype=np.int32))
        self._tri_med_out.append(np.full(n, -1 if mi[1] is None else mi[1], dtype=np.int32))
        self._nverts += verts.shape[0]
        self._ntris += n
        self._frozen = None

    def add_sphere(self, center, radius, material, emit=None, two_sided=False, n_samples=1, reverse_orientation=False):
        """Analytic sphere (pbrgpu_sphere; CPU twin host/AnalyticSphere.h - the reference's own Sphere is a stub).
        emit: RGB Lemit -> a DiffuseAreaLight over the sphere, sampled by solid angle."""
        light = -1
        if emit is not None:
            light = len(self.lights)
            self.lights.append(A.Light(type=A.LIGHT_AREA_SPHERE, tri=len(self.spheres), two_sided=int(two_sided),
                                       n_samples=n_samples, L=_f3(emit)))
        self.spheres.append(A.Sphere(center=_f3(center), radius=float(radius), material=int(material), light=light,
                                     reverse_orientation=int(reverse_orientation)))
        return len(self.spheres) - 1

    def add_point_light(self, pos, intensity):
        self.lights.append(A.Light(type=A.LIGHT_POINT, tri=-1, n_samples=1, pos=_f3(pos), L=_f3(intensity)))

    def add_infinite_light(self, env, scale=(1, 1, 1), light_to_world=None, n_samples=1):
        from .transform import Transform
        t = light_to_world if light_to_world is not None else Transform()
        self.env = np.ascontiguousarray(env, dtype=F)
        L = A.Light(type=A.LIGHT_INFINITE, tri=-1, n_samples=n_samples, L=_f3(scale))
        L.light_to_world = (C.c_float * 16)(*[float(x) for x in t.m.reshape(-1)])
        L.world_to_light = (C.c_float * 16)(*[float(x) for x in t.minv.reshape(-1)])
        self.lights.append(L)

    def add_skybox_light(self, env, world_radius, n_samples=1):
        self.env = np.ascontiguousarray(env, dtype=F)
        self.lights.append(A.Light(type=A.LIGHT_SKYBOX, tri=-1, n_samples=n_samples, world_radius=world_radius))

    # ---- flattened arrays ----
    def _freeze(self):
        if self._frozen is None:
            cat = lambda parts, shape, dt: (np.ascontiguousarray(np.concatenate(parts), dtype=dt) if parts
                                            else np.zeros(shape, dtype=dt))  # a scene of analytic spheres only has no mesh
            self._frozen = dict(
                verts=cat(self._verts, (0, 3), F), indices=cat(self._tris, (0, 3), np.int32),
                tri_material=cat(self._tri_mat, (0,), np.int32), tri_light=cat(self._tri_light, (0,), np.int32),
            )
        return self._frozen

    @property
    def verts(self):
        return self._freeze()["verts"]

    @property
    def indices(self):
        return self._freeze()["indices"]

    @property
    def tri_material(self):
        return self._freeze()["tri_material"]

    @property
    def tri_light(self):
        return self._freeze()["tri_light"]

    @property
    def n_tris(self):
        return self._ntris

    def desc(self):
        """-> (pbrgpu_scene_desc, keepalive list)"""
        fz = self._freeze()
        mats = (A.Material * max(1, len(self.materials)))(*self.materials)
        lights = (A.Light * max(1, len(self.lights)))(*self.lights)
        d = A.SceneDesc()
        d.n_verts = fz["verts"].shape[0]
        d.n_tris = fz["indices"].shape[0]
        d.verts = fz["verts"].ctypes.data_as(C.POINTER(C.c_float))
        d.indices = fz["indices"].ctypes.data_as(C.POINTER(C.c_int32))
        d.tri_material = fz["tri_material"].ctypes.data_as(C.POINTER(C.c_int32))
        d.tri_light = fz["tri_light"].ctypes.data_as(C.POINTER(C.c_int32))
        d.n_materials = len(self.materials)
        d.n_lights = len(self.lights)
        d.materials = mats
        d.lights = lights
        keep = [fz, mats, lights]
        if self.spheres:
            sph = (A.Sphere * len(self.spheres))(*self.spheres)
            d.n_spheres = len(self.spheres)
            d.spheres = sph
            keep.append(sph)
        fl = (np.ascontiguousarray(np.concatenate(self._mesh_flags), dtype=np.uint8) if self._mesh_flags
              else np.zeros(0, dtype=np.uint8))
        if self.tri_flags is not None:  # caller-supplied bit0: reverseOrientation (identity transforms: bit1 = bit0)
            user = np.ascontiguousarray(self.tri_flags, dtype=np.uint8)
            assert user.shape[0] == d.n_tris
            fl = (fl | np.where(user & 1, 3, 0)).astype(np.uint8)
        if fl.any():
            d.tri_flags = fl.ctypes.data_as(C.POINTER(C.c_uint8))
            keep.append(fl)
        if (fl & 4).any():
            nrm = np.ascontiguousarray(np.concatenate(self._normals), dtype=F)
            d.normals = nrm.ctypes.data_as(C.POINTER(C.c_float))
            keep.append(nrm)
        if (fl & 8).any():
            uv = np.ascontiguousarray(np.concatenate(self._uvs), dtype=F)
            d.uvs = uv.ctypes.data_as(C.POINTER(C.c_float))
            keep.append(uv)
        if (fl & 16).any():
            tg = np.ascontiguousarray(np.concatenate(self._tangents), dtype=F)
            d.tangents = tg.ctypes.data_as(C.POINTER(C.c_float))
            keep.append(tg)
        if self.env is not None:
            d.env_height, d.env_width = self.env.shape[0], self.env.shape[1]
            d.env_texels = self.env.ctypes.data_as(C.POINTER(C.c_float))
            keep.append(self.env)
        if self.textures:
            tex = (A.Texture * len(self.textures))(*[t for t, _ in self.textures])
            d.n_textures = len(self.textures)
            d.textures = tex
            keep += [tex, self.textures]
        if self.media:
            media = (A.Medium * len(self.media))(*self.media)
            mi = np.ascontiguousarray(np.concatenate(self._tri_med_in), dtype=np.int32)
            mo = np.ascontiguousarray(np.concatenate(self._tri_med_out), dtype=np.int32)
            d.n_media = len(self.media)
            d.media = media
            d.tri_medium_inside = mi.ctypes.data_as(C.POINTER(C.c_int32))
            d.tri_medium_outside = mo.ctypes.data_as(C.POINTER(C.c_int32))
            keep += [media, mi, mo]
        return d, keep


# ----------------------------------------------------------------------------------------------
# geometry helpers
# ----------------------------------------------------------------------------------------------
def uv_sphere(center, radius_fn, n_lat, n_lon):
    """(n_lat+1)*(n_lon+1) vertices (seam duplicated), 2*n_lat*n_lon triangles (pole triangles degenerate)."""
    th = (np.arange(n_lat + 1, dtype=np.float64) / n_lat) * np.pi
    ph = (np.arange(n_lon + 1, dtype=np.float64) / n_lon) * 2 * np.pi
    T, P = np.meshgrid(th, ph, indexing="ij")
    r = radius_fn(T, P) if callable(radius_fn) else np.full_like(T, radius_fn)
    x = r * np.sin(T) * np.cos(P)
    y = r * np.cos(T)
    z = r * np.sin(T) * np.sin(P)
    verts = np.stack([x, y, z], -1).reshape(-1, 3) + np.asarray(center, dtype=np.float64)
    i, j = np.meshgrid(np.arange(n_lat), np.arange(n_lon), indexing="ij")
    a = (i * (n_lon + 1) + j).reshape(-1)
    b = a + 1
    c = a + (n_lon + 1)
    d = c + 1
    # outward-facing winding for (theta down from +y, phi around y)
    tris = np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], 1).reshape(-1, 3)
    return verts.astype(F), tris.astype(np.int32)


def cornell_walls(scene, white, red, blue, length=5.0):
    # Main/main.cpp:288-311 vertex table; Translate(-L/2) applied as the TriangleMesh ctor would (x + (-2.5) in fp32)
    L = length
    P = np.array([
        [0, 0, L], [L, 0, L], [0, 0, 0], [L, 0, L], [L, 0, 0], [0, 0, 0],            # floor
        [0, L, L], [0, L, 0], [L, L, L], [L, L, L], [0, L, 0], [L, L, 0],            # ceiling
        [0, 0, 0], [L, 0, 0], [L, L, 0], [0, 0, 0], [L, L, 0], [0, L, 0],            # back wall
        [0, 0, 0], [0, L, L], [0, 0, L], [0, 0, 0], [0, L, 0], [0, L, L],            # right wall (x=0)
        [L, 0, 0], [L, L, L], [L, 0, L], [L, 0, 0], [L, L, 0], [L, L, L],            # left wall (x=L)
    ], dtype=F)
    P = (P + F(-0.5 * L)).astype(F)
    tris = np.arange(30, dtype=np.int32).reshape(10, 3)
    mats = np.array([white] * 6 + [red] * 2 + [blue] * 2, dtype=np.int32)  # main.cpp:317-323 (i==6,7 red; 8,9 blue)
    scene.add_mesh(P, tris, mats)


def cornell_area_light(scene, white, lemit=5.0, half=1.4, y=2.45, n_samples=5):
    # Main/main.cpp:353-375
    h = half
    P = np.array([[-h, 0, h], [-h, 0, -h], [h, 0, h], [h, 0, h], [-h, 0, -h], [h, 0, -h]], dtype=F)
    P[:, 1] = F(0.0) + F(y)
    scene.add_mesh(P, np.arange(6, dtype=np.int32).reshape(2, 3), white, emit=[lemit] * 3, two_sided=False,
                   n_samples=n_samples)


def cornell(sphere_tess=48, floor_mirror=False):
    """BASELINE config 1: Cornell box + area light + one Matte and one Mirror tessellated sphere."""
    s = SceneData()
    white = s.add_material(matte([1, 1, 1]))
    red = s.add_material(matte([0.9, 0.1, 0.17]))
    blue = s.add_material(matte([0.14, 0.21, 0.87]))
    mir = s.add_material(mirror([1, 1, 1]))
    green = s.add_material(matte([0.0, 1.0, 0.0]))
    cornell_walls(s, white, red, blue)
    if floor_mirror:
        s._tri_mat[-1][0:2] = mir
    cornell_area_light(s, white)
    if sphere_tess > 0:
        v, t = uv_sphere([-1.1, -1.6, -0.6], 0.9, sphere_tess, 2 * sphere_tess)
        s.add_mesh(v, t, green)
        v, t = uv_sphere([1.0, -1.6, 0.7], 0.9, sphere_tess, 2 * sphere_tess)
        s.add_mesh(v, t, mir)
    return s


def cornell_analytic(light_sphere=True, floor_mirror=False):
    """The Cornell box of `cornell()` with ANALYTIC spheres (pbrgpu_sphere / host/AnalyticSphere.h) instead of tessellated
    ones: a Matte, a Mirror, a Glass and a Plastic sphere, and (light_sphere) a small emissive sphere next to the ceiling
    light.  The reference's own Sphere is a stub, so the CPU side of the comparison is the twin."""
    s = SceneData()
    white = s.add_material(matte([0.73, 0.73, 0.73]))
    red = s.add_material(matte([0.9, 0.1, 0.17]))
    blue = s.add_material(matte([0.14, 0.21, 0.87]))
    cornell_walls(s, white, red, blue)
    cornell_area_light(s, white)
    m_matte = s.add_material(matte([0.2, 0.7, 0.3], sigma=20.0))
    m_mirror = s.add_material(mirror([0.9, 0.9, 0.9]))
    m_glass = s.add_material(glass([0.98, 0.98, 0.98], [0.98, 0.98, 0.98], 1.5))
    m_plastic = s.add_material(plastic([0.35, 0.12, 0.48], [0.65, 0.88, 0.52], 0.1))
    s.add_sphere([-1.2, -1.6, -0.6], 0.9, m_matte)
    s.add_sphere([1.1, -1.7, -1.0], 0.8, m_mirror)
    s.add_sphere([0.2, -1.9, 1.0], 0.6, m_glass)
    s.add_sphere([-0.4, 0.3, -1.4], 0.55, m_plastic)
    if light_sphere:
        s.add_sphere([1.4, 1.3, 0.4], 0.3, white, emit=[9.0, 7.0, 4.0], n_samples=2)
    return s


def cornell_media(kind="smoke", sphere_tess=24, sigma_a=(0.5, 0.5, 0.5), sigma_s=(4.4, 4.4, 4.4), g=-0.5):
    """Cornell box with a participating medium, the set-ups Main/main.cpp:242-243,346-348 sketches
    (HomogeneousMedium(0.5, 4.4, -0.5), MediumInterface(&medium, nullptr) on a closed mesh):
      "smoke"  the medium fills a sphere whose surface has NO material (pure medium boundary)
      "glass"  the medium fills a glass sphere (main.cpp:348)
      "fog"    no boundary at all: every primitive has an empty MediumInterface, so a ray keeps the medium it was
               born in - render it with an orthographic camera created with `media=0` (global fog)
    Returns (SceneData, medium id)."""
    s = SceneData()
    white = s.add_material(matte([1, 1, 1]))
    red = s.add_material(matte([0.9, 0.1, 0.17]))
    blue = s.add_material(matte([0.14, 0.21, 0.87]))
    mir = s.add_material(mirror([1, 1, 1]))
    gls = s.add_material(glass([0.98] * 3, [0.98] * 3, 1.5))
    cornell_walls(s, white, red, blue)
    cornell_area_light(s, white)
    med = s.add_medium(sigma_a, sigma_s, g)
    if kind == "fog":
        v, t = uv_sphere([0.9, -1.6, 0.4], 0.9, sphere_tess, 2 * sphere_tess)
        s.add_mesh(v, t, mir)
        return s, med
    v, t = uv_sphere([-0.9, -1.2, 0.2], 1.1, sphere_tess, 2 * sphere_tess)
    s.add_mesh(v, t, None if kind == "smoke" else gls, medium_interface=(med, None))
    v, t = uv_sphere([1.2, -1.7, 0.9], 0.8, sphere_tess, 2 * sphere_tess)
    s.add_mesh(v, t, mir)
    return s, med


def dragon_class(tess=467, with_point_light=False, block=8192, volume=None):
    """BASELINE config 2 (and 4): ~872k-triangle bumpy sphere, half Matte(0,1,0) / half Mirror in
    triangle-index blocks, inside the Cornell walls under the Cornell area light.
    volume = "smoke" | "glass": the mesh becomes the boundary of HomogeneousMedium(0.5, 4.4, -0.5) (Main/main.cpp:242-243,
    346-348) - without material, or as glass - for the VolPath integrator."""
    s = SceneData()
    white = s.add_material(matte([1, 1, 1]))
    red = s.add_material(matte([0.9, 0.1, 0.17]))
    blue = s.add_material(matte([0.14, 0.21, 0.87]))
    green = s.add_material(matte([0.0, 1.0, 0.0]))
    mir = s.add_material(mirror([1, 1, 1]))
    cornell_walls(s, white, red, blue)
    cornell_area_light(s, white)

    def radius(T, P):
        return 1.4 * (1 + 0.08 * np.sin(9 * T) * np.cos(11 * P) + 0.03 * np.sin(37 * P + 5 * T))

    v, t = uv_sphere([0.0, -0.6, 0.0], radius, tess, 2 * tess)
    mats = np.where((np.arange(t.shape[0]) // block) % 2 == 0, green, mir).astype(np.int32)
    if volume is None:
        s.add_mesh(v, t, mats)
    else:
        med = s.add_medium([0.5] * 3, [4.4] * 3, -0.5)
        gls = s.add_material(glass([0.98] * 3, [0.98] * 3, 1.5))
        s.add_mesh(v, t, None if volume == "smoke" else gls, medium_interface=(med, None))
    if with_point_light:
        s.add_point_light([0.0, 2.0, 0.0], [50.0, 50.0, 50.0])
    return s


def procedural_env(width=2048, height=1024, seed=7, sun=1.0e4):
    """Smooth sky gradient + one bright sun disc + seeded value noise."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height, dtype=np.float64) + 0.5) / height
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    U, V = np.meshgrid(u, v)
    sky = np.stack([0.35 + 0.4 * (1 - V), 0.5 + 0.4 * (1 - V), 0.7 + 0.5 * (1 - V)], -1)
    ground = np.stack([0.25 + 0 * V, 0.22 + 0 * V, 0.2 + 0 * V], -1)
    img = np.where((V < 0.5)[..., None], sky, ground)
    coarse = rng.random((height // 32 + 2, width // 32 + 2))
    yi = (np.arange(height) / 32.0)
    xi = (np.arange(width) / 32.0)
    y0 = yi.astype(int)
    x0 = xi.astype(int)
    fy = (yi - y0)[:, None]
    fx = (xi - x0)[None, :]
    n = (coarse[y0][:, x0] * (1 - fy) * (1 - fx) + coarse[y0 + 1][:, x0] * fy * (1 - fx) +
         coarse[y0][:, x0 + 1] * (1 - fy) * fx + coarse[y0 + 1][:, x0 + 1] * fy * fx)
    img = img * (0.8 + 0.4 * n[..., None])
    d2 = ((U - 0.3) * 2.0) ** 2 + (V - 0.25) ** 2
    img = img + sun * np.exp(-d2 / (2 * 0.004 ** 2))[..., None] * np.array([1.0, 0.95, 0.85])
    return np.ascontiguousarray(img, dtype=F)


def bsdf_stress(sphere_tess=64, env=None, skybox=False):
    """BASELINE config 3: 3x3 grid of Glass / Metal / Plastic spheres over a Matte ground under an env map."""
    from .transform import Transform
    s = SceneData()
    ground = s.add_material(matte([0.5, 0.5, 0.5]))
    mats = [
        s.add_material(glass([0.98] * 3, [0.98] * 3, 1.5, 0.0, 0.0, False)),
        s.add_material(glass([0.98] * 3, [0.98] * 3, 1.5, 0.1, 0.1, False)),
        s.add_material(metal([0.2, 0.2, 0.8], [0.11] * 3, 0.15, 0.15, False)),
        s.add_material(plastic([0.35, 0.12, 0.48], [0.65, 0.88, 0.52], 0.1, True)),
    ]
    g = 6.0
    P = np.array([[-g, -1, -g], [g, -1, -g], [g, -1, g], [-g, -1, g]], dtype=F)
    s.add_mesh(P, np.array([[0, 2, 1], [0, 3, 2]], dtype=np.int32), ground)
    k = 0
    for iz in range(3):
        for ix in range(3):
            v, t = uv_sphere([(ix - 1) * 2.2, -0.1, (iz - 1) * 2.2], 0.9, sphere_tess, 2 * sphere_tess)
            s.add_mesh(v, t, mats[k % len(mats)])
            k += 1
    if env is None:
        env = procedural_env(512, 256)
    if skybox:
        s.add_skybox_light(env, world_radius=20.0, n_samples=10)
    else:
        s.add_infinite_light(env, (1, 1, 1), Transform(), n_samples=10)
    return s
