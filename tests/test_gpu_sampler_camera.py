"""GPU parity: Halton sampler values and primary rays are BIT-EXACT against the compiled reference.
Reference: Sampler/Halton.cpp:61-92, Sampler/LowDiscrepancy.cpp:211-232,2300-2316, Sampler/Sampler.cpp:10-21,
Camera/Perspective.cpp:44-80, Camera/Orthographic.cpp:7-25."""
import numpy as np
import pytest

from pysicalbasedraytracer_b200 import _abi as A
from pysicalbasedraytracer_b200 import scenes
from tests.util import Pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny(api, ref, gpu_ctx):
    return Pair(api, ref, scenes.cornell(sphere_tess=0), 512, 512, 16)


def test_perms_bit_exact(tiny, ref):
    want = ref.perms()
    got = tiny.scene.dump_perms(len(want))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("region,s0,s1", [((0, 0, 8, 8), 0, 4), ((3, 5, 4, 6), 0, 16), ((500, 250, 512, 256), 7, 9),
                                          ((127, 127, 130, 130), 100000, 100002)])
def test_sampler_stream_bit_exact(tiny, region, s0, s1):
    n_dims = 90  # > 5 + 8*10: every dimension a depth-10 path can touch
    want = tiny.rsmp.dump(region, s0, s1, n_dims)
    got = tiny.scene.dump_sampler(region, s0, s1, n_dims)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_sampler_other_bounds(api, ref, gpu_ctx):
    # small images change baseScales/Exponents/stride (Halton.cpp:40-57); non-square too
    for (W, H, spp) in [(64, 64, 4), (20, 9, 3), (1920, 1080, 2)]:
        p = Pair(api, ref, scenes.cornell(sphere_tess=0), W, H, spp)
        region = (W - 3, H - 2, W, H)
        want = p.rsmp.dump(region, 0, spp, 24)
        got = p.scene.dump_sampler(region, 0, spp, 24)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (W, H)


def test_primary_rays_bit_exact(tiny, ref):
    for region in [(0, 0, 16, 16), (3, 5, 4, 6), (496, 496, 512, 512)]:
        want = ref.primary_rays(tiny.rcam, tiny.rsmp, region, 0, 4)
        got = tiny.scene.dump_primary_rays(region, 0, 4)
        # o.xyz, d.xyz bit-exact; tMax = inf; time is dead state on this path (we store 0)
        assert np.array_equal(got[..., [0, 1, 2, 4, 5, 6]].view(np.uint32), want[..., [0, 1, 2, 4, 5, 6]].view(np.uint32))
        assert np.all(np.isinf(got[..., 3]))


def test_primary_rays_golden_survey(tiny):
    # pixel (3,5) s0 of the 512x512 perspective camera at (0,0,5)->(0,0,0)
    got = tiny.scene.dump_primary_rays((3, 5, 4, 6), 0, 1).view(np.uint32).reshape(-1)
    assert [hex(v) for v in got[[0, 1, 2, 4, 5, 6]]] == ['0x0', '0x0', '0x40a00000', '0x3f139459', '0x3f122bae', '0xbf15a267']


def test_orthographic_and_nonsquare_rays(api, ref, gpu_ctx):
    for cam_type, W, H in [(A.CAMERA_ORTHOGRAPHIC, 256, 256), (A.CAMERA_PERSPECTIVE, 320, 180),
                           (A.CAMERA_ORTHOGRAPHIC, 100, 160)]:
        p = Pair(api, ref, scenes.cornell(sphere_tess=0), W, H, 4, eye=(0.3, 0.2, 5.0), look=(0.1, -0.2, 0.0),
                 cam_type=cam_type)
        region = (0, 0, W, 2)
        want = ref.primary_rays(p.rcam, p.rsmp, region, 0, 2)
        got = p.scene.dump_primary_rays(region, 0, 2)
        assert np.array_equal(got[..., [0, 1, 2, 4, 5, 6]].view(np.uint32), want[..., [0, 1, 2, 4, 5, 6]].view(np.uint32))


def test_sobol_sampler_bit_exact_vs_cpu_twin(api, ref, gpu_ctx):
    """Sobol': PARITY UNPINNED by the reference (it has the matrices, no sampler).  The pin is the CPU twin
    host/SobolSampler.h - a GlobalSampler subclass running inside the reference's own Sampler machinery over the
    reference's own tables - against the GPU sampler: values and primary rays bit-exact."""
    for (W, H, spp) in [(512, 512, 16), (1920, 1080, 4), (20, 9, 3)]:
        p = Pair(api, ref, scenes.cornell(sphere_tess=0), W, H, spp, sampler_type=A.SAMPLER_SOBOL)
        region = (W - 3, H - 2, W, H)
        want = p.rsmp.dump(region, 0, spp, 64)
        got = p.scene.dump_sampler(region, 0, spp, 64)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (W, H)
        assert 0.0 <= got.min() and got.max() < 1.0
        wr = ref.primary_rays(p.rcam, p.rsmp, region, 0, 2)
        gr = p.scene.dump_primary_rays(region, 0, 2)
        assert np.array_equal(gr[..., [0, 1, 2, 4, 5, 6]].view(np.uint32), wr[..., [0, 1, 2, 4, 5, 6]].view(np.uint32))
    # (0,2)-sequence property: 16 consecutive samples of one pixel stratify the pixel into a 4x4 grid
    p = Pair(api, ref, scenes.cornell(sphere_tess=0), 64, 64, 16, sampler_type=A.SAMPLER_SOBOL)
    d = p.scene.dump_sampler((10, 20, 11, 21), 0, 16, 2)[0]
    cells = set((int(u * 4), int(v * 4)) for u, v in d)
    assert len(cells) == 16
