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


def test_primary_hits(cornell, ref):
    region = (0, 0, 256, 256)
    want = ref.primary_hits(cornell.rscene, cornell.rcam, cornell.rsmp, region, 0, 2)
    got = cornell.scene.dump_primary_hits(region, 0, 2)
    frac = _compare_hits(got, want)
    print("primary-hit mismatch fraction", frac)
    assert frac < MAX_MISMATCH


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


def test_random_closest_hits(cornell):
    rays = _random_rays(200_000, 1)
    want = cornell.rscene.trace(rays)
    got = cornell.scene.trace(rays)
    frac = _compare_hits(got, want)
    print("random closest mismatch fraction", frac)
    assert frac < MAX_MISMATCH


def test_random_any_hits(cornell):
    rays = _random_rays(200_000, 2)
    rays[:, 3] = np.random.default_rng(3).uniform(0.05, 3.0, rays.shape[0]).astype(np.float32)
    want = cornell.rscene.trace(rays, any_hit=True)
    got = cornell.scene.trace(rays, any_hit=True)
    frac = (want != got).mean()
    print("random any-hit mismatch fraction", frac)
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


def test_dragon_class_primary_hits(api, ref, gpu_ctx):
    # reduced tessellation of the BASELINE config-2 mesh (full size is exercised by bench.py)
    p = Pair(api, ref, scenes.dragon_class(tess=96), 192, 108, 2)
    region = (0, 0, 192, 108)
    want = ref.primary_hits(p.rscene, p.rcam, p.rsmp, region, 0, 2)
    got = p.scene.dump_primary_hits(region, 0, 2)
    frac = _compare_hits(got, want)
    print("dragon-class primary-hit mismatch fraction", frac, "bvh", p.scene.info())
    assert frac < MAX_MISMATCH
