"""CPU tests (no GPU) of the host-side logic: camera-matrix construction (bit-exact against the compiled
reference and the golden fixtures), scene generators, sample-range sharding, and the world_size-2
film reduce over gloo."""
import os
import socket

import numpy as np
import pytest

from pysicalbasedraytracer_b200 import _abi as A
from pysicalbasedraytracer_b200 import scenes, shard
from tests.util import golden

KAT = golden("kat.npz")


@pytest.mark.parametrize("name,cam_type,W,H,eye,look", [
    ("persp512", A.CAMERA_PERSPECTIVE, 512, 512, (0.0, 0.0, 5.0), (0.0, 0.0, 0.0)),
    ("persp320x180", A.CAMERA_PERSPECTIVE, 320, 180, (0.3, 0.2, 5.0), (0.1, -0.2, 0.0)),
    ("ortho100x160", A.CAMERA_ORTHOGRAPHIC, 100, 160, (0.3, 0.2, 5.0), (0.1, -0.2, 0.0))])
def test_camera_matrices_bit_exact_vs_golden(api, name, cam_type, W, H, eye, look):
    c2w = api.Inverse(api.LookAt(eye, look, (0.0, 1.0, 0.0)))
    cam = (api.CreatePerspectiveCamera if cam_type == A.CAMERA_PERSPECTIVE else api.CreateOrthographicCamera)(W, H, c2w)
    d = cam.desc()
    assert np.array_equal(np.array(d.raster_to_camera[:], dtype=np.float32).view(np.uint32), KAT[f"cam_{name}_r2c"].view(np.uint32))
    assert np.array_equal(np.array(d.camera_to_world[:], dtype=np.float32).view(np.uint32), KAT[f"cam_{name}_c2w"].view(np.uint32))


def test_camera_matrices_bit_exact_vs_reference(api, ref):
    rng = np.random.default_rng(3)
    for _ in range(12):
        eye = rng.uniform(-6, 6, 3)
        look = rng.uniform(-1, 1, 3)
        W, H = int(rng.integers(16, 2000)), int(rng.integers(16, 1200))
        for cam_type in (A.CAMERA_PERSPECTIVE, A.CAMERA_ORTHOGRAPHIC):
            rc = ref.RefCamera(cam_type, W, H, eye, look, (0, 1, 0))
            want = rc.desc()
            c2w = api.Inverse(api.LookAt(eye, look, (0.0, 1.0, 0.0)))
            cam = (api.CreatePerspectiveCamera if cam_type == A.CAMERA_PERSPECTIVE else api.CreateOrthographicCamera)(W, H, c2w)
            got = cam.desc()
            for f in ("raster_to_camera", "camera_to_world"):
                g = np.array(getattr(got, f)[:], dtype=np.float32)
                w = np.array(getattr(want, f)[:], dtype=np.float32)
                assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), (f, W, H, cam_type)


def test_scene_generators_are_deterministic_and_sized():
    a, b = scenes.cornell(sphere_tess=8), scenes.cornell(sphere_tess=8)
    assert np.array_equal(a.verts, b.verts) and np.array_equal(a.indices, b.indices)
    assert a.n_tris == 12 + 2 * (2 * 8 * 16)
    d = scenes.dragon_class(tess=467)
    assert d.n_tris == 872368  # BASELINE configs[1]: ~870k triangles
    assert len(d.lights) == 2 and d.tri_light.max() == 1
    e = scenes.procedural_env(64, 32)
    assert e.shape == (32, 64, 3) and e.dtype == np.float32 and np.isfinite(e).all() and e.max() > 1.0


def test_shard_samples_partition():
    for (s0, s1, world) in [(0, 1024, 8), (0, 16, 8), (3, 10, 4), (0, 5, 8), (0, 4096, 2), (7, 8, 1)]:
        parts = [shard.shard_samples(s0, s1, r, world) for r in range(world)]
        assert parts[0][0] == s0 and parts[-1][1] == s1
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == s1 - s0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("wrap", [A.WRAP_REPEAT, A.WRAP_BLACK, A.WRAP_CLAMP])
def test_lanczos_resampling_bit_exact_vs_reference_mipmap(api, ref, wrap):
    """Images whose sides are not powers of two: the library's host resampler (what pbrgpu_scene_create applies) against level 0
    of the reference's own MIPMap (Texture/MIPMap.h:93-144), read back through ImageTexture::Evaluate at the texel centres with a
    zero footprint (bilinear weights 1, 0, 0, 0 -> the texel itself)."""
    rng = np.random.default_rng(17)
    for (h, w) in [(24, 50), (5, 3), (33, 64), (100, 37)]:
        img = rng.uniform(0.0, 2.0, (h, w, 3)).astype(np.float32)
        got = api.resample_pow2(img, wrap)
        H2, W2 = got.shape[:2]
        assert (W2 & (W2 - 1)) == 0 and (H2 & (H2 - 1)) == 0 and W2 >= w and H2 >= h and W2 < 2 * w and H2 < 2 * h
        tex = A.Texture(width=w, height=h, texels=img.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)),
                        do_trilinear=0, max_anisotropy=8.0, wrap=wrap, su=1.0, sv=1.0, du=0.0, dv=0.0)
        rows = np.zeros((H2 * W2, 8), dtype=np.float32)
        uu, vv = np.meshgrid((np.arange(W2, dtype=np.float32) + np.float32(0.5)) / np.float32(W2),
                             (np.arange(H2, dtype=np.float32) + np.float32(0.5)) / np.float32(H2))
        rows[:, 0], rows[:, 1] = uu.reshape(-1), vv.reshape(-1)
        want = ref.texture_kat((tex, img), rows).reshape(H2, W2, 3)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (h, w, wrap)
    same = api.resample_pow2(np.ones((8, 16, 3), dtype=np.float32))
    assert same.shape == (8, 16, 3)


def test_shard_samples_partition_properties(api):
    """pbrgpu_shard_samples (the split pbrgpu_render applies on a multi-GPU context): the shards tile [s0, s1) without gaps
    or overlap, in rank order, sizes differ by at most one; ranks beyond the sample count get empty ranges."""
    for s0, s1 in [(0, 1024), (0, 4096), (7, 8), (5, 5), (100, 111), (0, 3), (16, 16 + 128)]:
        for world in (1, 2, 3, 4, 8):
            cuts = [api.shard_samples(s0, s1, r, world) for r in range(world)]
            assert cuts[0][0] == s0 and cuts[-1][1] == s1
            sizes = [hi - lo for lo, hi in cuts]
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1 and sum(sizes) == s1 - s0
            assert sizes == sorted(sizes, reverse=True)  # the first (n % world) ranks take the extra sample
            for r, (lo, hi) in enumerate(cuts):
                assert (lo, hi) == shard.shard_samples(s0, s1, r, world)
    with pytest.raises(api.PbrGpuError):
        api.shard_samples(0, 8, 2, 2)
    with pytest.raises(api.PbrGpuError):
        api.shard_samples(9, 8, 0, 2)


def test_multi_gpu_entry_points_fail_loudly_without_a_gpu(api):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.PbrGpuError, match="no CUDA device"):
        api.Context(devices=[0, 1])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W, spp = 6, 5, 11
    lo, hi = shard.shard_samples(0, spp, rank, world)
    # stand-in for the renderer: the "radiance" of (pixel, sample) is a closed-form function of both, so the
    # reduced film has a known answer
    film = torch.zeros((H, W, 4), dtype=torch.float32)
    pix = torch.arange(H * W, dtype=torch.float32).reshape(H, W)
    for s in range(lo, hi):
        film[..., 0] += pix + s
        film[..., 1] += 2 * pix - s
        film[..., 2] += 0.5 * s
        film[..., 3] += 1
    shard.reduce_film(film, dst=0)
    if rank == 0:
        q.put(film.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_film_reduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    film = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    H, W, spp = 6, 5, 11
    pix = np.arange(H * W, dtype=np.float32).reshape(H, W)
    s = np.arange(spp, dtype=np.float32)
    assert np.allclose(film[..., 0], spp * pix + s.sum())
    assert np.allclose(film[..., 1], 2 * spp * pix - s.sum())
    assert np.allclose(film[..., 2], 0.5 * s.sum())
    assert np.all(film[..., 3] == spp)
    import torch
    mean = shard.resolve_mean(torch.from_numpy(film)).numpy()
    assert np.allclose(mean[..., 0], pix + s.mean())


def test_3d_mesh_loader_matches_reference_plyinfo(api, ref, tmp_path):
    """Scene ingestion: the reference's private .3d text format (Shape/plyRead.h:17-47)."""
    rng = np.random.default_rng(9)
    nv, nt = 257, 400
    verts = rng.normal(size=(nv, 3)) * 0.07
    tris = rng.integers(0, nv, size=(nt, 3))
    path = tmp_path / "mesh.3d"
    with open(path, "w") as f:
        f.write(f"vertex {nv}\nface {nt}\n")
        for v in verts:
            f.write(f"{v[0]:.9g} {v[1]:.7f} {v[2]:.6e}\n")
        for t in tris:
            f.write(f"3 {t[0]} {t[1]} {t[2]}\n")
    gv, gt = api.load_3d(path)
    wv, wt = ref.ply_read(path)
    assert gv.shape == (nv, 3) and gt.shape == (nt, 3)
    assert np.array_equal(gv.view(np.uint32), wv.view(np.uint32))  # same parse, same x20 float multiply
    assert np.array_equal(gt, wt)
    with pytest.raises(api.PbrGpuError):
        api.load_3d(tmp_path / "missing.3d")
    # the loaded soup feeds straight into the scene description
    s = scenes.SceneData()
    m = s.add_material(scenes.matte([0.5, 0.5, 0.5]))
    s.add_mesh(gv, gt, m)
    assert s.n_tris == nt


def test_media_scene_description_and_golden_fixture_match_reference(ref):
    """SceneData.add_medium / medium_interface / material-less meshes flatten into the descriptor the C-ABI documents, and the
    committed VolPath fixtures are what the compiled reference's VolPathIntegrator::Li renders today."""
    from tests.util import golden
    data, med = scenes.cornell_media("smoke", sphere_tess=12)
    d, keep = data.desc()
    assert d.n_media == 1 and med == 0
    assert abs(d.media[0].sigma_s[1] - 4.4) < 1e-6 and abs(d.media[0].g + 0.5) < 1e-6
    tri_mat = np.ctypeslib.as_array(d.tri_material, (d.n_tris,))
    m_in = np.ctypeslib.as_array(d.tri_medium_inside, (d.n_tris,))
    m_out = np.ctypeslib.as_array(d.tri_medium_outside, (d.n_tris,))
    boundary = tri_mat == -1
    assert 0 < boundary.sum() < d.n_tris
    assert (m_in[boundary] == 0).all() and (m_in[~boundary] == -1).all() and (m_out == -1).all()
    cam = api_camera_desc_default()
    assert cam.medium_plus1 == 0
    # fixture == reference (CPU, 64x64x4 spp)
    g = golden("volpath64_smoke.npz")
    rsc = ref.RefScene(data)
    rc = ref.RefCamera(A.CAMERA_PERSPECTIVE, 64, 64, (0.0, 0.0, 5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    img, st = ref.render(rsc, rc, 4, (0, 0, 64, 64), integrator=A.INTEGRATOR_VOLPATH, max_depth=5, rr_threshold=0.8)
    assert np.array_equal((img / np.float32(4)).astype(np.float32), g["rgb_mean"])
    assert int(st.rays_closest) == int(g["rays_closest"]) and int(st.rays_shadow) == 0


def api_camera_desc_default():
    from pysicalbasedraytracer_b200 import api
    cam = api.CreateOrthographicCamera(16, 16, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0))), media=0)
    return cam.desc()  # Camera::medium is recorded but, as in the reference, not delivered to the rays


def _write_hdr(path, rgbe, mode):
    """rgbe uint8 [H,W,4] -> Radiance file.  mode: "rle" (new run-length scanlines), "flat" (plain quadruples)."""
    h, w, _ = rgbe.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# written by the test\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        if mode == "flat":
            f.write(rgbe.tobytes())
            return
        for j in range(h):
            f.write(bytes([2, 2, w >> 8, w & 255]))
            for k in range(4):
                row = rgbe[j, :, k]
                i = 0
                while i < w:
                    run = 1
                    while i + run < w and run < 127 and row[i + run] == row[i]:
                        run += 1
                    if run >= 3:
                        f.write(bytes([128 + run, int(row[i])]))
                        i += run
                    else:
                        lit = 1
                        while i + lit < w and lit < 128 and not (i + lit + 2 < w and row[i + lit] == row[i + lit + 1] == row[i + lit + 2]):
                            lit += 1
                        f.write(bytes([lit]) + row[i:i + lit].tobytes())
                        i += lit


def test_hdr_loader_matches_reference_stb(api, ref, tmp_path):
    """pbrgpu_load_hdr vs the reference's stbi_loadf (what InfiniteAreaLight / SkyBoxLight read): texels bit-identical,
    run-length coded and flat files, with and without the vertical flip SkyBoxLight switches on."""
    rng = np.random.default_rng(5)
    for name, (w, h, mode) in {"rle": (64, 32, "rle"), "flat_small": (5, 7, "flat"), "flat_wide": (40, 9, "flat"),
                               "rle_odd": (37, 11, "rle")}.items():
        rgbe = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        rgbe[:, :, 3] = rng.integers(118, 140, (h, w))
        rgbe[::3, : w // 2, :] = rgbe[0, 0, :]  # long runs
        rgbe[1, 1, 3] = 0                        # exponent 0 -> exactly black
        if mode == "flat" and w >= 8:
            rgbe[0, 0, 0] = 7                    # a flat file must not start like a run-length scanline header
        path = tmp_path / f"{name}.hdr"
        _write_hdr(path, rgbe, mode)
        for flip in (False, True):
            got = api.load_hdr(path, flip_vertically=flip)
            want = ref.load_hdr(path, flip=flip)
            assert got.shape == (h, w, 3) and want.shape[2] == 3
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, flip)
        assert (got[h - 2 if flip else 1, 1] == 0).all()
    with pytest.raises(api.PbrGpuError):
        api.load_hdr(tmp_path / "missing.hdr")
    bad = tmp_path / "bad.hdr"
    bad.write_bytes(b"P6\n2 2\n255\n" + bytes(12))
    with pytest.raises(api.PbrGpuError):
        api.load_hdr(bad)
    trunc = tmp_path / "trunc.hdr"
    trunc.write_bytes((tmp_path / "rle.hdr").read_bytes()[:200])
    with pytest.raises(api.PbrGpuError):
        api.load_hdr(trunc)
