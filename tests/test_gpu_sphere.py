"""Analytic sphere primitive (include/pbrgpu.h pbrgpu_sphere, csrc/sphere.cuh).

PARITY UNPINNED BY THE REFERENCE: its own PBR::Sphere is a stub (Shape/Sphere.cpp:8-37: Intersect returns `discriminant > 0`
without a hit record, IntersectP returns false, Sample returns an empty Interaction), so no reference output exists for these
values.  The oracle side here is the CPU twin host/AnalyticSphere.h - a subclass of the reference's Shape that implements the
quadratic Shape/README.md section 3 documents - running inside the reference's own BVHAccel / GeometricPrimitive /
DiffuseAreaLight / integrators.  Pins: known-answer rays with closed-form answers, twin == GPU (t bit-exact; radiance 1e-3)."""
import numpy as np
import pytest

from pysicalbasedraytracer_b200 import _abi as A
from pysicalbasedraytracer_b200 import scenes
from tests.test_gpu_render import REL_TOL, assert_diverged_is_discrete
from tests.util import Pair, rel_err

pytestmark = pytest.mark.gpu


def _one_sphere(center=(0.5, -0.25, 1.0), radius=1.5, emit=None):
    s = scenes.SceneData()
    m = s.add_material(scenes.matte([0.5, 0.5, 0.5]))
    s.add_sphere(center, radius, m, emit=emit)
    if emit is None:
        s.add_point_light([0, 4, 0], [10, 10, 10])
    return s


def _rays(n, seed, scale=4.0, tmax=np.inf):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = rng.uniform(-scale, scale, (n, 3))
    d = rng.normal(size=(n, 3))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3] = tmax
    return rays


def test_known_answer_rays(api, gpu_ctx):
    """Closed-form answers: rays along the axes through the centre of a sphere of radius 1.5 at (0.5, -0.25, 1)."""
    c, r = np.array([0.5, -0.25, 1.0], dtype=np.float32), 1.5
    p = Pair(api, None, _one_sphere(c, r), 16, 16, 1)
    rays = np.zeros((6, 8), dtype=np.float32)
    rays[:, 3] = np.inf
    # from outside along +x: enters at distance 5 - r; from the centre: leaves at r; tMax short of the sphere: miss;
    # tangent line (discriminant == 0): miss, like the reference stub's `discriminant > 0`; pointing away: miss
    rays[0, 0:3] = c + [-5, 0, 0]; rays[0, 4:7] = [1, 0, 0]
    rays[1, 0:3] = c;              rays[1, 4:7] = [0, 0, 1]
    rays[2, 0:3] = c + [-5, 0, 0]; rays[2, 4:7] = [1, 0, 0]; rays[2, 3] = 3.0
    rays[3, 0:3] = c + [-5, r, 0]; rays[3, 4:7] = [1, 0, 0]
    rays[4, 0:3] = c + [-5, 0, 0]; rays[4, 4:7] = [-1, 0, 0]
    rays[5, 0:3] = c + [0, 0, 4];  rays[5, 4:7] = [0, 0, -2]   # un-normalised direction: t is in units of |d|
    for via in (False, True):   # helper loop and the production kernel k_trace<.., SPH>
        got = p.scene.trace(rays, via_wave_kernel=via)
        t = got[:, 1].view(np.float32)
        assert list(got[:, 0]) == [0, 0, -1, -1, -1, 0]
        assert t[0] == np.float32(3.5) and t[1] == np.float32(1.5) and t[5] == np.float32(1.25)
        occl = p.scene.trace(rays, any_hit=True, via_wave_kernel=via)
        assert list(occl) == [1, 1, 0, 0, 0, 1]


@pytest.fixture(scope="module")
def box(api, ref, gpu_ctx):
    return Pair(api, ref, scenes.cornell_analytic(), 96, 96, 4)


@pytest.mark.parametrize("kernel", ["k_trace", "helper_loop"])
def test_intersect_and_intersectp_match_the_twin(box, ref, kernel):
    """AnalyticSphere::Intersect / IntersectP inside the reference's BVHAccel vs the GPU: primitive ids identical, t bit-exact
    (the quadratic is evaluated in double on both sides)."""
    via = kernel == "k_trace"
    rays = _rays(200_000, 11, scale=2.4)
    want, got = box.rscene.trace(rays), box.scene.trace(rays, via_wave_kernel=via)
    n_tris = box.scene.data.n_tris
    assert (want[:, 0] >= n_tris).mean() > 0.1            # a good share of the rays ends on a sphere
    mism = want[:, 0] != got[:, 0]
    print(kernel, "closest-hit mismatch fraction", mism.mean())
    assert mism.mean() < 1e-4
    assert np.array_equal(want[~mism, 1], got[~mism, 1])  # t, bit for bit
    rays[:, 3] = np.random.default_rng(12).uniform(0.05, 3.0, rays.shape[0])
    wa, ga = box.rscene.trace(rays, any_hit=True), box.scene.trace(rays, any_hit=True, via_wave_kernel=via)
    assert (wa != ga).mean() < 1e-4
    region = (0, 0, 96, 96)
    wh = ref.primary_hits(box.rscene, box.rcam, box.rsmp, region, 0, 2)
    gh = box.scene.dump_wave_hits(region, 0, 2) if via else box.scene.dump_primary_hits(region, 0, 2)
    assert (wh[..., 0] != gh[..., 0]).mean() < 1e-4
    same = wh[..., 0] == gh[..., 0]
    assert np.array_equal(wh[..., 1][same], gh[..., 1][same])


def test_sample_and_pdf_match_the_twin(api, ref, gpu_ctx):
    """AnalyticSphere::Sample(ref, u) / Pdf(ref, wi) through DiffuseAreaLight::Sample_Li / Pdf_Li: reference points outside
    (uniform cone) and inside (uniform area) the emissive sphere."""
    c, r = np.array([0.5, -0.25, 1.0], dtype=np.float32), 1.5
    data = _one_sphere(c, r, emit=[3.0, 2.0, 1.0])
    p = Pair(api, ref, data, 16, 16, 1)
    rng = np.random.default_rng(5)
    n = 4096
    rows = np.zeros((n, 16), dtype=np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dist = np.where(np.arange(n) % 4 == 0, rng.uniform(0.05, 0.95, n), rng.uniform(1.05, 6.0, n)) * r  # a quarter inside
    rows[:, 0:3] = c + d * dist[:, None]
    rows[:, 3:6] = 1e-6
    nrm = rng.normal(size=(n, 3))
    rows[:, 6:9] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    rows[:, 9:11] = rng.uniform(0, 1, (n, 2))
    q = (c - rows[:, 0:3]) + rng.normal(size=(n, 3)) * 0.6 * r                                           # roughly towards the sphere
    rows[:, 11:14] = q / np.linalg.norm(q, axis=1, keepdims=True)
    want, got = p.rscene.light_eval(0, rows), p.scene.light_eval(0, rows)
    inside = dist < r
    for name, sl, tol in (("Li", slice(0, 3), 0), ("pdf", slice(3, 4), 2e-5), ("wi", slice(4, 7), 2e-6), ("p1", slice(7, 10), 2e-6),
                          ("n1", slice(10, 13), 2e-6), ("Pdf_Li", slice(13, 14), 2e-5)):
        a, b = want[:, sl].astype(np.float64), got[:, sl].astype(np.float64)
        err = np.abs(a - b) / np.maximum(np.abs(a), 1.0 if name in ("wi", "p1", "n1") else 1e-3)
        print(name, "max rel err outside", err[~inside].max(), "inside", err[inside].max())
        assert err.max() <= tol, name
    assert (want[:, 3] > 0).mean() > 0.95 and (want[:, 13] > 0).mean() > 0.3   # both branches produce samples / densities


@pytest.mark.parametrize("integ,depth", [(A.INTEGRATOR_PATH, 1), (A.INTEGRATOR_PATH, 5), (A.INTEGRATOR_WHITTED, 4),
                                          (A.INTEGRATOR_DIRECT_ALL, 3), (A.INTEGRATOR_VOLPATH, 4)])
def test_renders_match_the_twin(box, ref, integ, depth):
    """All five materials on analytic spheres, a triangle area light and an emissive sphere, through the reference's own
    integrators (CPU twin) and the GPU wavefront, at matched samples."""
    kw = dict(integrator=integ, max_depth=depth)
    if integ in (A.INTEGRATOR_PATH, A.INTEGRATOR_VOLPATH):
        kw["rr_threshold"] = 0.8
    want, rst = ref.render(box.rscene, box.rcam, box.spp, (0, 0, box.W, box.H), s1=2, **kw)
    got, _, st = box.scene.render(s1=2, **kw)
    err = rel_err(got / 2, want / 2)
    ok = (err <= REL_TOL).mean()
    print(f"integrator {integ} depth {depth}: pixels within {REL_TOL}: {ok:.5f}, median err {np.median(err):.2e}, "
          f"rays gpu {st.rays_closest}+{st.rays_shadow} twin {rst.rays_closest}+{rst.rays_shadow}")
    assert ok >= 0.999
    assert_diverged_is_discrete(err)
    assert abs(int(st.rays_closest) - int(rst.rays_closest)) <= 2e-3 * rst.rays_closest + 8


def test_cpp_dropin_flattens_analytic_spheres(api, ref, gpu_ctx):
    """host/GpuIntegrator.h recognises AnalyticSphere shapes in the reference's object graph."""
    data = scenes.cornell_analytic()
    rsc = ref.RefScene(data)
    rc = ref.RefCamera(A.CAMERA_PERSPECTIVE, 64, 64, (0, 0, 5), (0, 0, 0), (0, 1, 0))
    rgba, rgb_sum, _ = ref.gpu_dropin_render(rsc, rc, 4, 64, 64, 4, 0.8, "uniform")
    want, _ = ref.render(rsc, rc, 4, (0, 0, 64, 64), integrator=A.INTEGRATOR_PATH, max_depth=4, rr_threshold=0.8)
    assert (rel_err(rgb_sum / 4, want / 4) <= REL_TOL).mean() >= 0.999
    assert rgba[..., 3].min() == 255


def test_sphere_error_behaviour(api, gpu_ctx):
    s = scenes.SceneData()
    m = s.add_material(scenes.matte([0.5, 0.5, 0.5]))
    s.add_sphere([0, 0, 0], -1.0, m)
    with pytest.raises(api.PbrGpuError, match="radius must be positive"):
        api.Scene(s)
    s = scenes.SceneData()
    s.add_material(scenes.matte([0.5, 0.5, 0.5]))
    s.add_sphere([0, 0, 0], 1.0, 3)
    with pytest.raises(api.PbrGpuError, match="sphere material out of range"):
        api.Scene(s)
    s = scenes.SceneData()
    tex = s.add_texture(np.ones((4, 4, 3), dtype=np.float32))
    s.add_sphere([0, 0, 0], 1.0, s.add_material(scenes.matte([0.5, 0.5, 0.5], kd_texture=tex)))
    with pytest.raises(api.PbrGpuError, match="image textures on analytic spheres"):
        api.Scene(s)
