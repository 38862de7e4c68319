"""GPU parity: BVH traversal + watertight triangle test against the reference's
BVHAccel::Intersect / IntersectP (Accelerator/BVHAccel.cpp:285-366, Shape/Triangle.cpp:62-350).
Hit primitive ids must be identical; disagreements (ties at equal t, float-epsilon box misses in
the reference's un-widened slab test) are counted and must stay below 1e-4."""
import numpy as np
import pytest

from pysicalbasedraytracer_b200 import scenes
from tests.util import Pair

pytestmark = pytest.mark.gpu
MAX_MISMATCH = 1e-4


@pytest.fixture(scope="module")
def cornell(api, ref, gpu_ctx):
    return Pair(api, ref, scenes.cornell(sphere_tess=32), 256, 256, 4)


def _compare_hits(got, want):
    prim_g, prim_w = got[..., 0].reshape(-1), want[..., 0].reshape(-1)
    t_g, t_w = got[..., 1].reshape(-1).view(np.float32), want[..., 1].reshape(-1).view(np.float32)
    mism = prim_g != prim_w
    frac = mism.mean()
    same = ~mism
    # where the primitive agrees, t must be bit-identical (same triangle arithmetic)
    assert np.array_equal(t_g[same].view(np.uint32), t_w[same].view(np.uint32))
    return frac


# Every hit-id check runs twice: on the PRODUCTION kernel of a render (k_wave_begin -> k_raygen -> k_trace: persistent warps,
# parked triangle groups, mid-warp refetch, shared-memory hit record; PBRGPU_DUMP_WAVE_HITS / PBRGPU_TRACE_VIA_WAVE_KERNEL) and
# on the one-thread-per-ray helper loop (PBRGPU_DUMP_PRIMARY_HITS / plain pbrgpu_trace_rays) as a second, independent pin.
KERNELS = ["k_trace", "helper_loop"]


def _primary(scene, region, s0, s1, kernel):
    return scene.dump_wave_hits(region, s0, s1) if kernel == "k_trace" else scene.dump_primary_hits(region, s0, s1)


@pytest.mark.parametrize("kernel", KERNELS)
def test_primary_hits(cornell, ref, kernel):
    region = (0, 0, 256, 256)
    want = ref.primary_hits(cornell.rscene, cornell.rcam, cornell.rsmp, region, 0, 2)
    got = _primary(cornell.scene, region, 0, 2, kernel)
    frac = _compare_hits(got, want)
    print(kernel, "primary-hit mismatch fraction", frac)
    assert frac < MAX_MISMATCH


def test_wave_kernel_equals_helper_loop(cornell):
    """The two device paths share the node / triangle steps; their answers must be identical."""
    region = (13, 40, 256, 211)
    a = cornell.scene.dump_wave_hits(region, 1, 4)
    b = cornell.scene.dump_primary_hits(region, 1, 4)
    assert np.array_equal(a, b)


def _random_rays(n, seed, tmax=np.inf):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-2.4, 2.4, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = tmax
    rays[:, 4:7] = d
    return rays


@pytest.mark.parametrize("kernel", KERNELS)
def test_random_closest_hits(cornell, kernel):
    # k_trace: the rays go through its path queue AND its MIS queue in one launch (the call fails if the two disagree)
    rays = _random_rays(200_000, 1)
    want = cornell.rscene.trace(rays)
    got = cornell.scene.trace(rays, via_wave_kernel=kernel == "k_trace")
    frac = _compare_hits(got, want)
    print(kernel, "random closest mismatch fraction", frac)
    assert frac < MAX_MISMATCH


@pytest.mark.parametrize("kernel", KERNELS)
def test_random_any_hits(cornell, kernel):
    # k_trace: through its shadow queue (any-hit early-out, contribution added when unoccluded)
    rays = _random_rays(200_000, 2)
    rays[:, 3] = np.random.default_rng(3).uniform(0.05, 3.0, rays.shape[0]).astype(np.float32)
    want = cornell.rscene.trace(rays, any_hit=True)
    got = cornell.scene.trace(rays, any_hit=True, via_wave_kernel=kernel == "k_trace")
    frac = (want != got).mean()
    print(kernel, "random any-hit mismatch fraction", frac)
    assert frac < MAX_MISMATCH
    assert 0.05 < want.mean() < 0.95  # the test exercises both outcomes


def test_axis_aligned_and_degenerate_rays(cornell):
    # rays with zero direction components (inf/NaN slabs) and rays starting on geometry
    rays = np.zeros((6, 8), dtype=np.float32)
    rays[:, 3] = np.inf
    for k in range(6):
        rays[k, 4 + k % 3] = 1.0 if k < 3 else -1.0
        rays[k, 0:3] = (0.3, 0.1, -0.2)
    want = cornell.rscene.trace(rays)
    got = cornell.scene.trace(rays)
    assert np.array_equal(got[:, 0], want[:, 0])
    assert np.array_equal(got[:, 1], want[:, 1])


@pytest.mark.parametrize("kernel", KERNELS)
def test_near_axis_parallel_rays(api, ref, gpu_ctx, kernel):
    """Directions whose smallest component is 2^-6 .. 2^-16 of the largest: the half-precision box test (trace.cuh
    cw_node_test_h2) loses range there and hands over to the fp32 test at a ratio of 2^11; hits must not change on either
    side of the switch, for rays from outside the mesh and for rays that start inside its bounds."""
    p = Pair(api, ref, scenes.dragon_class(tess=96), 64, 64, 1)
    rng = np.random.default_rng(11)
    n = 60_000
    rays = _random_rays(n, 5)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    small = rng.integers(0, 3, n)
    ratio = np.exp2(-rng.uniform(6, 16, n)).astype(np.float32)
    d[np.arange(n), small] = np.abs(d).max(axis=1) * ratio * rng.choice([-1.0, 1.0], n).astype(np.float32)
    both = rng.random(n) < 0.3  # two small components: the ray runs along one axis
    d[both, (small[both] + 1) % 3] = np.abs(d[both]).max(axis=1) * ratio[both] * 0.5
    rays[:, 4:7] = d
    rays[n // 2:, 0:3] *= 0.3  # origins inside the mesh's bounds
    want = p.rscene.trace(rays)
    got = p.scene.trace(rays, via_wave_kernel=kernel == "k_trace")
    assert (want[:, 0] >= 0).mean() > 0.5
    assert _compare_hits(got, want) < MAX_MISMATCH
    rays[:, 3] = rng.uniform(0.05, 3.0, n).astype(np.float32)
    occ = p.rscene.trace(rays, any_hit=True)
    assert (occ != p.scene.trace(rays, any_hit=True, via_wave_kernel=kernel == "k_trace")).mean() < MAX_MISMATCH
    assert 0.02 < occ.mean() < 0.98


@pytest.mark.parametrize("scale,shift", [(1e-3, (0.0, 0.0, 0.0)), (1e3, (0.0, 0.0, 0.0)), (1.0, (3000.0, -2000.0, 1500.0)),
                                         (0.05, (40.0, 40.0, -40.0))])
def test_scaled_and_translated_scenes(api, ref, gpu_ctx, scale, shift):
    """The box test works relative to each node (pivot at the node entry, per-node power-of-two scale): hits must not depend on
    the size of the scene's coordinates or on its distance from the origin (large |origin - o| / grid unit is where the fp32
    part of the error bound matters).  A flat, strongly anisotropic mesh (node exponents far apart per axis) is included."""
    rng = np.random.default_rng(21)
    s = scenes.SceneData()
    m = s.add_material(scenes.matte([0.5, 0.5, 0.5]))
    sh = np.asarray(shift, dtype=np.float64)
    verts, tris = scenes.uv_sphere((0.0, 0.0, 0.0), lambda T, P: 1.0 + 0.05 * np.sin(9 * T) * np.cos(7 * P), 96, 192)
    s.add_mesh((verts.astype(np.float64) * scale + sh).astype(np.float32), tris, m)
    # a 400 x 1 grid of thin quads: long in x, flat in y
    gx = np.linspace(-2.0, 2.0, 401)
    strip = np.stack([np.stack([gx, 0.001 * np.sin(40 * gx), np.full_like(gx, z)], axis=1) for z in (-1.6, -1.55)]).reshape(-1, 3)
    idx = np.arange(400)
    stris = np.concatenate([np.stack([idx, idx + 1, idx + 401], axis=1), np.stack([idx + 1, idx + 402, idx + 401], axis=1)]).astype(np.int32)
    s.add_mesh((strip * scale + sh).astype(np.float32), stris, m)
    s.add_point_light(list(sh + np.array([0.0, 3.0, 0.0]) * scale), [10, 10, 10])
    p = Pair(api, ref, s, 32, 32, 1, eye=tuple(sh + np.array([0.0, 0.0, 4.0]) * scale), look=tuple(sh))
    n = 60_000
    rays = _random_rays(n, 31)
    rays[:, 0:3] = (rays[:, 0:3].astype(np.float64) * scale + sh).astype(np.float32)
    aim = (rng.normal(size=(n, 3)) * 0.6 * scale + sh) - rays[:, 0:3]  # aimed at the mesh, so that most rays hit
    rays[:, 4:7] = (aim / np.linalg.norm(aim, axis=1, keepdims=True)).astype(np.float32)
    want = p.rscene.trace(rays)
    assert (want[:, 0] >= 0).mean() > 0.3
    for via in (False, True):
        assert _compare_hits(p.scene.trace(rays, via_wave_kernel=via), want) < MAX_MISMATCH


@pytest.mark.parametrize("kernel", KERNELS)
def test_dragon_class_primary_hits(api, ref, gpu_ctx, kernel):
    # reduced tessellation of the BASELINE config-2 mesh (the full 872 k-triangle mesh: test_gpu_render.py::test_full_size_*)
    p = Pair(api, ref, scenes.dragon_class(tess=96), 192, 108, 2)
    region = (0, 0, 192, 108)
    want = ref.primary_hits(p.rscene, p.rcam, p.rsmp, region, 0, 2)
    got = _primary(p.scene, region, 0, 2, kernel)
    frac = _compare_hits(got, want)
    print(kernel, "dragon-class primary-hit mismatch fraction", frac, "bvh", p.scene.info())
    assert frac < MAX_MISMATCH
    rays = _random_rays(100_000, 7)
    assert _compare_hits(p.scene.trace(rays, via_wave_kernel=kernel == "k_trace"), p.rscene.trace(rays)) < MAX_MISMATCH


def test_tiny_scenes_and_edge_cases(api, ref, gpu_ctx):
    """Degenerate sizes: a one-triangle scene (the BVH root is a single leaf), 2- and 3-triangle scenes, rays that
    start on geometry, zero-length and NaN directions."""
    for n_tris in (1, 2, 3, 9):
        s = scenes.SceneData()
        m = s.add_material(scenes.matte([0.5, 0.5, 0.5]))
        rng = np.random.default_rng(n_tris)
        verts = rng.uniform(-1, 1, (3 * n_tris, 3)).astype(np.float32)
        s.add_mesh(verts, np.arange(3 * n_tris, dtype=np.int32).reshape(-1, 3), m)
        s.add_point_light([0, 3, 0], [10, 10, 10])
        p = Pair(api, ref, s, 32, 32, 1, eye=(0.0, 0.0, 4.0))
        rays = _random_rays(20000, 100 + n_tris)
        rays[:, 0:3] *= 0.6
        want, got = p.rscene.trace(rays), p.scene.trace(rays)
        assert np.array_equal(got[:, 0], want[:, 0]) and np.array_equal(got[:, 1], want[:, 1]), n_tris
        assert (want[:, 0] >= 0).any()
        img, _, st = p.scene.render(max_depth=2)
        wimg, _ = ref.render(p.rscene, p.rcam, 1, (0, 0, 32, 32), max_depth=2)
        assert (np.abs(img - wimg).max(-1) <= 1e-3 * np.maximum(np.abs(wimg).max(-1), 1e-3)).mean() >= 0.995
    # degenerate rays must not hang or crash: zero direction, NaN direction, origin exactly on a triangle vertex.
    # (Their *answer* is unspecified: with d = 0 the reference's own triangle test runs on NaNs, every comparison is
    # false and it reports a hit with t = NaN, Triangle.cpp:84-145; only termination is asserted here.)
    rays = np.zeros((3, 8), dtype=np.float32)
    rays[:, 3] = np.inf
    rays[1, 4:7] = np.nan
    rays[2, 0:3] = verts[0]
    rays[2, 4:7] = (0, 0, 1)
    out = p.scene.trace(rays)
    assert out.shape == (3, 4)
    assert p.scene.trace(rays, any_hit=True).shape == (3,)


def test_c_abi_error_behaviour(api, gpu_ctx):
    """Errors are reported through the int status + pbrgpu_last_error(), never by crashing."""
    import ctypes as C
    from pysicalbasedraytracer_b200 import _abi as A
    s = scenes.SceneData()
    m = s.add_material(scenes.matte([0.5, 0.5, 0.5]))
    s.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32), np.array([[0, 1, 2]], dtype=np.int32), m)
    sc = api.Scene(s)
    with pytest.raises(api.PbrGpuError, match="camera / sampler not set"):
        sc.render(region=(0, 0, 4, 4), s1=1)
    sc.set_camera(api.CreatePerspectiveCamera(16, 16, api.Inverse(api.LookAt((0, 0, 3), (0, 0, 0), (0, 1, 0)))))
    sc.set_sampler(api.HaltonSampler(4, (0, 0, 16, 16)))
    with pytest.raises(api.PbrGpuError, match="empty pixel region"):
        sc.render(region=(4, 4, 4, 8))
    with pytest.raises(api.PbrGpuError, match="bad sample range"):
        sc.render(s0=3, s1=1)
    with pytest.raises(api.PbrGpuError, match="unknown integrator"):
        sc.render(integrator=17)
    img, _, st = sc.render(s0=2, s1=2)  # empty sample range: nothing rendered, zeros returned
    assert st.camera_samples == 0 and not img.any()
    img, _, _ = sc.render(region=(5, 7, 6, 8), s1=4)  # a single pixel
    assert img.shape == (1, 1, 3) and np.isfinite(img).all()
    bad = scenes.SceneData()
    bad.add_material(scenes.matte([1, 1, 1]))
    bad.add_mesh(np.zeros((3, 3), dtype=np.float32), np.array([[0, 1, 7]], dtype=np.int32), 0)
    with pytest.raises(api.PbrGpuError, match="vertex index out of range"):
        api.Scene(bad)
    d = A.SamplerDesc(type=A.SAMPLER_SOBOL, samples_per_pixel=4)
    with pytest.raises(api.PbrGpuError, match="Sobol needs"):
        sc.set_sampler(d)


def test_bvh_build_is_independent_of_the_thread_count(api, gpu_ctx, monkeypatch):
    """The host builder (csrc/bvh_build.cpp) builds the top of the binary SAH tree on the calling thread and the subtrees below
    on a pool of threads; nodes, leaf order - and therefore every ray query - must not depend on how many threads there were."""
    data = scenes.dragon_class(tess=96)  # 36 k triangles: above the builder's single-thread threshold
    rays = _random_rays(20000, 3)
    got = []
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("PBRGPU_BUILD_THREADS", threads)
        sc = api.Scene(data, gpu_ctx)
        got.append((sc.info(), sc.trace(rays)))
        sc.close()
    for info, hits in got[1:]:
        assert info == got[0][0] and info["bvh_hash"] != 0
        assert np.array_equal(hits, got[0][1])
