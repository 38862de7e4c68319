"""Generates the committed golden fixtures from the COMPILED, UNMODIFIED reference (oracle/_ref,
built by oracle/build_ref.sh from /root/reference).  Run where /root/reference exists:

    python tests/golden/make_golden.py

Outputs (small .npz files next to this script) pin the CPU restatement (oracle/restate.c) and the GPU
path on machines where neither /root/reference nor oracle/_ref is available."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refharness as R  # noqa: E402
from pysicalbasedraytracer_b200 import _abi as A  # noqa: E402
from pysicalbasedraytracer_b200 import scenes  # noqa: E402

EYE, LOOK, UP = (0.0, 0.0, 5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)


def fnv1a64(a):
    h = 0xcbf29ce484222325
    for v in np.asarray(a, dtype=np.uint16).tolist():
        h ^= v
        h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def cam_arrays(rc):
    d = rc.desc()
    return np.array(d.raster_to_camera[:], dtype=np.float32), np.array(d.camera_to_world[:], dtype=np.float32)


def main():
    out = {}
    # ---- sampler ----
    perms = R.perms()
    out["perms_len"] = np.int64(len(perms))
    out["perms_fnv1a64"] = np.uint64(fnv1a64(perms))
    out["perms_head"] = perms[:64]
    for name, (W, H, spp) in {"b512": (512, 512, 16), "b1080": (1920, 1080, 4), "b20x9": (20, 9, 3)}.items():
        s = R.RefSampler(spp, (0, 0, W, H))
        region = (3, 5, 6, 7) if W > 8 else (0, 0, 2, 2)
        out[f"halton_{name}_params"] = np.array(s.params(), dtype=np.int32)
        out[f"halton_{name}_region"] = np.array(region, dtype=np.int32)
        out[f"halton_{name}_dump"] = s.dump(region, 0, 3, 40)
        out[f"halton_{name}_index"] = np.array([s.index(region[0], region[1], k) for k in (0, 1, 2, 1000)], dtype=np.int64)
    out["radinv"] = np.array([R.lib().ref_radical_inverse(b, a) for b, a in
                              [(0, 1), (1, 5), (0, 0xdeadbeef), (7, 123456789), (85, 31104 * 1024 + 77)]], dtype=np.float32)
    out["scrambled"] = np.array([R.lib().ref_scrambled_radical_inverse(b, a) for b, a in
                                 [(2, 7), (5, 12345), (40, 99999999), (86, 31104 * 1023 + 5)]], dtype=np.float32)
    # ---- cameras / primary rays ----
    for name, (ct, W, H, eye, look) in {"persp512": (A.CAMERA_PERSPECTIVE, 512, 512, EYE, LOOK),
                                        "persp320x180": (A.CAMERA_PERSPECTIVE, 320, 180, (0.3, 0.2, 5.0), (0.1, -0.2, 0.0)),
                                        "ortho100x160": (A.CAMERA_ORTHOGRAPHIC, 100, 160, (0.3, 0.2, 5.0), (0.1, -0.2, 0.0))}.items():
        rc = R.RefCamera(ct, W, H, eye, look, UP)
        rs = R.RefSampler(4, (0, 0, W, H))
        r2c, c2w = cam_arrays(rc)
        out[f"cam_{name}_r2c"], out[f"cam_{name}_c2w"] = r2c, c2w
        out[f"cam_{name}_rays"] = R.primary_rays(rc, rs, (3, 5, 7, 7), 0, 2)
    # ---- primary hits + random rays on a small Cornell scene ----
    data = scenes.cornell(sphere_tess=8)
    rsc = R.RefScene(data)
    rc = R.RefCamera(A.CAMERA_PERSPECTIVE, 64, 64, EYE, LOOK, UP)
    rs = R.RefSampler(4, (0, 0, 64, 64))
    out["hits_cornell8_primary"] = R.primary_hits(rsc, rc, rs, (0, 0, 64, 64), 0, 1)
    rng = np.random.default_rng(11)
    rays = np.zeros((4096, 8), dtype=np.float32)
    rays[:, 0:3] = rng.uniform(-2.4, 2.4, (4096, 3))
    d = rng.normal(size=(4096, 3))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3] = np.inf
    out["rays_random"] = rays
    out["hits_cornell8_random"] = rsc.trace(rays)
    rays2 = rays.copy()
    rays2[:, 3] = rng.uniform(0.05, 3.0, 4096)
    out["rays_random_any"] = rays2
    out["hits_cornell8_any"] = rsc.trace(rays2, any_hit=True)
    # ---- radiance: Cornell path depth 5 (smoke fixture) and depth 1 (direct lighting) ----
    data = scenes.cornell(sphere_tess=12)
    rsc = R.RefScene(data)
    img, st = R.render(rsc, rc, 4, (0, 0, 64, 64), integrator=A.INTEGRATOR_PATH, max_depth=5, rr_threshold=0.8)
    np.savez_compressed(os.path.join(HERE, "cornell64_path5.npz"), rgb_mean=(img / np.float32(4)).astype(np.float32),
                        rays_closest=np.int64(st.rays_closest), rays_shadow=np.int64(st.rays_shadow))
    img, st = R.render(rsc, rc, 4, (0, 0, 64, 64), integrator=A.INTEGRATOR_PATH, max_depth=1, rr_threshold=0.8)
    np.savez_compressed(os.path.join(HERE, "cornell64_direct.npz"), rgb_mean=(img / np.float32(4)).astype(np.float32),
                        rays_closest=np.int64(st.rays_closest), rays_shadow=np.int64(st.rays_shadow))
    # materials under the importance-sampled env map (direct lighting)
    env = scenes.procedural_env(128, 64)
    data = scenes.bsdf_stress(sphere_tess=12, env=env)
    rsc = R.RefScene(data)
    rc2 = R.RefCamera(A.CAMERA_PERSPECTIVE, 96, 64, (0.0, 3.0, 7.0), (0.0, -0.3, 0.0), UP)
    img, st = R.render(rsc, rc2, 2, (0, 0, 96, 64), integrator=A.INTEGRATOR_PATH, max_depth=1)
    np.savez_compressed(os.path.join(HERE, "stress96x64_direct.npz"), rgb_mean=(img / np.float32(2)).astype(np.float32),
                        rays_closest=np.int64(st.rays_closest), rays_shadow=np.int64(st.rays_shadow))
    # VolPathIntegrator: smoke sphere (medium behind a material-less boundary) + glass sphere filled with the medium
    for kind in ("smoke", "glass"):
        data, _ = scenes.cornell_media(kind, sphere_tess=12)
        rsc = R.RefScene(data)
        img, st = R.render(rsc, rc, 4, (0, 0, 64, 64), integrator=A.INTEGRATOR_VOLPATH, max_depth=5, rr_threshold=0.8)
        np.savez_compressed(os.path.join(HERE, f"volpath64_{kind}.npz"), rgb_mean=(img / np.float32(4)).astype(np.float32),
                            rays_closest=np.int64(st.rays_closest), rays_shadow=np.int64(st.rays_shadow))
    # ---- participating media known answers: HomogeneousMedium::Sample / Tr, HenyeyGreenstein::Sample_p / p ----
    rows = np.zeros((512, 8), dtype=np.float32)
    d = rng.normal(size=(512, 3))
    rows[:, 0:3] = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.2, 3.0, (512, 1))
    rows[:, 3] = np.where(rng.uniform(size=512) < 0.2, np.inf, rng.uniform(0.01, 4.0, 512))
    rows[:, 4:6] = rng.uniform(0, 1, (512, 2))
    rows[:4, 4:6] = [[0.0, 0.0], [0.99999994, 0.99999994], [0.33333334, 0.5], [0.6666667, 1e-7]]
    out["medium_kat_in"] = rows
    for tag, (sa, ss, g) in {"a": ((0.5, 0.5, 0.5), (4.4, 4.4, 4.4), -0.5), "b": ((0.05, 0.04, 0.03), (0.25, 0.3, 0.35), 0.3),
                             "c": ((0.0, 1.0, 0.2), (0.0, 0.0, 2.0), 0.0)}.items():
        out[f"medium_kat_{tag}_params"] = np.array(list(sa) + list(ss) + [g], dtype=np.float32)
        out[f"medium_kat_{tag}_out"] = R.medium_kat(sa, ss, g, rows)
    rows = np.zeros((512, 8), dtype=np.float32)
    wo = rng.normal(size=(512, 3)); wi = rng.normal(size=(512, 3))
    rows[:, 0:3] = wo / np.linalg.norm(wo, axis=1, keepdims=True)
    rows[:, 3:5] = rng.uniform(0, 1, (512, 2))
    rows[:, 5:8] = wi / np.linalg.norm(wi, axis=1, keepdims=True)
    rows[:3, 0:3] = [[0, 0, 1], [1, 0, 0], [0, -1, 0]]
    out["hg_kat_in"] = rows
    for g in (-0.5, 0.0, 0.3, 0.0005, 0.95):
        out[f"hg_kat_out_{g}"] = R.hg_kat(g, rows)
    # ---- film tone map ----
    rgb = rng.uniform(0, 1.5, (512, 3)).astype(np.float32)
    rgb[:8] = [[0, 0, 0], [1, 1, 1], [0.0031308, 0.0031308, 0.0031308], [2, 0.5, 0.1], [1e-4, 1e-3, 1e-2], [0.5, 0.5, 0.5],
               [0.2, 0.7, 0.1], [10, 10, 10]]
    out["tonemap_rgb"] = rgb
    out["tonemap_rgba8"] = R.tonemap(rgb)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
