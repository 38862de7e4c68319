"""Multi-GPU behind the C-ABI (include/pbrgpu.h "multi-GPU"): scene replicated per GPU, sample range sharded,
ONE ncclReduce(sum) of the fp32 RGBA films to rank 0.  Both ways in:
  * pbrgpu_init_multi     - one process drives N devices (what host/GpuIntegrator.h::setDevices uses)
  * pbrgpu_comm_attach    - one process per device (torchrun); exercised here with two spawned processes
The N-GPU film must equal the 1-GPU film of the same sample range up to fp32 summation order (<= 2e-6 relative).
Tests that need two devices skip on a one-GPU box (the driver's round-end box); run them with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest

from pysicalbasedraytracer_b200 import _abi as A
from pysicalbasedraytracer_b200 import scenes

pytestmark = pytest.mark.gpu
W = H = 96
SPP = 10
KW = dict(integrator=A.INTEGRATOR_PATH, max_depth=4, rr_threshold=0.8)


def _setup(api, ctx, data=None):
    sc = api.Scene(data or scenes.cornell(sphere_tess=12), ctx)
    sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
    sc.set_sampler(api.HaltonSampler(SPP, (0, 0, W, H)))
    return sc


def _rel(a, b):
    den = np.maximum(np.abs(b).max(-1), 1e-3)
    return float((np.abs(a - b).max(-1) / den).max())


def test_group_context_of_one_device_is_the_plain_path(api, gpu_ctx):
    ctx = api.Context(devices=[0])
    assert ctx.info() == dict(n_devices=1, rank=0, nranks=1, nccl_version=0) or ctx.info()["n_devices"] == 1
    a, rgba_a, st_a = _setup(api, ctx).render(want_rgba8=True, **KW)
    b, rgba_b, st_b = _setup(api, gpu_ctx).render(want_rgba8=True, **KW)
    assert np.array_equal(a, b) and np.array_equal(rgba_a, rgba_b)
    assert st_a.rays_closest == st_b.rays_closest and st_a.camera_samples == W * H * SPP
    ctx.close()


def _need(api, n):
    if api.device_count() < n:
        pytest.skip(f"needs {n} GPUs (run under `gpurun --gpus {n}`)")


@pytest.mark.parametrize("n_dev", [2, 4])
def test_group_context_shards_and_reduces(api, gpu_ctx, n_dev):
    _need(api, n_dev)
    ctx = api.Context(devices=list(range(n_dev)))
    info = ctx.info()
    assert info["n_devices"] == n_dev and info["nccl_version"] > 0
    multi = _setup(api, ctx)
    single = _setup(api, gpu_ctx)
    a, rgba_a, st_a = multi.render(want_rgba8=True, **KW)
    b, rgba_b, st_b = single.render(want_rgba8=True, **KW)
    err = _rel(a, b)
    print(f"{n_dev} devices vs 1: max rel err {err:.2e}, rays {st_a.rays_closest}+{st_a.rays_shadow} vs {st_b.rays_closest}+{st_b.rays_shadow}")
    assert err <= 2e-6                                   # only the summation order differs
    assert np.abs(rgba_a.astype(int) - rgba_b.astype(int)).max() <= 1
    assert st_a.camera_samples == st_b.camera_samples    # statistics are summed over the devices
    assert st_a.rays_closest == st_b.rays_closest and st_a.rays_shadow == st_b.rays_shadow
    # fewer samples than devices: some shards are empty and still take part in the reduce
    c, _, _ = multi.render(s0=3, s1=4, **KW)
    d, _, _ = single.render(s0=3, s1=4, **KW)
    assert np.array_equal(c, d)
    # PBRGPU_RENDER_LOCAL_ONLY: no sharding, bit-identical to the single-GPU context
    e, _, _ = multi.render(flags=A.RENDER_LOCAL_ONLY, **KW)
    assert np.array_equal(e, b)
    # the parity hooks of a group scene run on its first device
    assert np.array_equal(multi.dump_wave_hits((0, 0, W, H), 0, 1), single.dump_wave_hits((0, 0, W, H), 0, 1))
    # VolPath reads counters back inside the render: one host thread per device
    v, _, _ = multi.render(integrator=A.INTEGRATOR_VOLPATH, max_depth=4)
    w, _, _ = single.render(integrator=A.INTEGRATOR_VOLPATH, max_depth=4)
    assert _rel(v, w) <= 2e-6
    ctx.close()


def test_cpp_dropin_on_two_devices(api, ref, gpu_ctx):
    """host/GpuIntegrator.h::setDevices({0, 1}): the application's single Integrator::Render call on two GPUs."""
    _need(api, 2)
    data = scenes.cornell(sphere_tess=12)
    rsc = ref.RefScene(data)
    rc = ref.RefCamera(A.CAMERA_PERSPECTIVE, 64, 64, (0, 0, 5), (0, 0, 0), (0, 1, 0))
    rgba2, sum2, _ = ref.gpu_dropin_render(rsc, rc, 8, 64, 64, 5, 0.8, "uniform", n_devices=2)
    rgba1, sum1, _ = ref.gpu_dropin_render(rsc, rc, 8, 64, 64, 5, 0.8, "uniform", n_devices=1)
    assert _rel(sum2, sum1) <= 2e-6
    assert np.abs(rgba2.astype(int) - rgba1.astype(int)).max() <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_worker(rank, world, uid_q, out_q):
    # one process per GPU, no torch: the unique id travels through a multiprocessing queue (any launcher-side channel does)
    from pysicalbasedraytracer_b200 import api
    ctx = api.Context(rank)
    if rank == 0:
        uid = api.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    ctx.attach(world, rank, uid)
    sc = _setup(api, ctx)
    img, rgba, st = sc.render(want_rgba8=True, **KW)          # every rank: the same call
    if rank == 0:
        alone, _, _ = sc.render(flags=A.RENDER_LOCAL_ONLY, **KW)
        out_q.put((img, rgba, alone, int(st.camera_samples)))
    else:
        assert not img.any()                                  # only rank 0's buffers are written
        out_q.put(int(st.camera_samples))
    sc.close()
    ctx.close()


def test_process_per_gpu_attach(api, gpu_ctx):
    _need(api, 2)
    import multiprocessing as mp
    mpx = mp.get_context("spawn")
    world = 2
    uid_q, out0, out1 = mpx.Queue(), mpx.Queue(), mpx.Queue()
    procs = [mpx.Process(target=_rank_worker, args=(r, world, uid_q, out0 if r == 0 else out1)) for r in range(world)]
    for p in procs:
        p.start()
    img, rgba, alone, n0 = out0.get(timeout=300)
    n1 = out1.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert _rel(img, alone) <= 2e-6
    assert n0 + n1 == W * H * SPP and n0 == n1  # 10 samples -> 5 + 5
    assert rgba[..., 3].min() == 255
