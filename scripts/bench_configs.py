"""Throughput of the other BASELINE.json configs on one GPU (bench.py covers configs[1]); development aid.
Each config renders a bounded slice (spp samples of every pixel) of the quoted job."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

EYE, LOOK, UP = (0, 0, 5), (0, 0, 0), (0, 1, 0)


def run(name, data, W, H, spp_job, spp, integrator=A.INTEGRATOR_PATH, max_depth=8, rr=1.0, eye=EYE, look=LOOK, reps=3):
    sc = api.Scene(data)
    sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt(eye, look, UP))))
    sc.set_sampler(api.HaltonSampler(spp_job, (0, 0, W, H)))
    best = None
    out = np.zeros((H, W, 3), dtype=np.float32)
    for r in range(reps):
        _, _, st = sc.render(integrator=integrator, max_depth=max_depth, rr_threshold=rr, s0=0, s1=spp, out=out)
        if best is None or st.ms_total < best.ms_total:
            best = st
    rays = best.rays_closest + best.rays_shadow
    line = dict(config=name, tris=data.n_tris, resolution=[W, H], spp_slice=spp, ms=round(best.ms_total, 2),
                msamples_per_s=round(best.camera_samples / best.ms_total / 1e3, 1), mrays_per_s=round(rays / best.ms_total / 1e3, 1),
                rays_per_sample=round(rays / best.camera_samples, 2), mean_radiance=float(out.mean() / spp), bvh=sc.info())
    print(json.dumps(line), flush=True)
    sc.close()


if __name__ == "__main__":
    run("1 cornell 512x512 16spp path d10 rr0.8", scenes.cornell(sphere_tess=48), 512, 512, 16, 16, max_depth=10, rr=0.8)
    env = scenes.procedural_env(2048, 1024)
    run("3a bsdf-stress InfiniteAreaLight 1080p (16 of 1024 spp) path d8", scenes.bsdf_stress(sphere_tess=64, env=env), 1920, 1080, 1024, 16,
        eye=(0.0, 3.0, 7.0), look=(0.0, -0.3, 0.0))
    run("3b bsdf-stress SkyBoxLight 1080p (16 of 1024 spp) path d8", scenes.bsdf_stress(sphere_tess=64, env=env, skybox=True), 1920, 1080, 1024, 16,
        eye=(0.0, 3.0, 7.0), look=(0.0, -0.3, 0.0))
    d4 = scenes.dragon_class(tess=467, with_point_light=True)
    run("4a dragon+point+area Whitted d5 1080p 8spp", d4, 1920, 1080, 1024, 8, integrator=A.INTEGRATOR_WHITTED, max_depth=5)
    run("4b dragon+point+area Direct/UniformSampleOne d5 1080p 8spp", d4, 1920, 1080, 1024, 8, integrator=A.INTEGRATOR_DIRECT_ONE, max_depth=5)
    run("4c dragon+point+area Direct/UniformSampleAll d5 1080p 4spp", d4, 1920, 1080, 1024, 4, integrator=A.INTEGRATOR_DIRECT_ALL, max_depth=5)
    d5 = scenes.dragon_class(tess=467)
    d5.add_skybox_light(env, world_radius=20.0, n_samples=1)
    run("5 dragon + SkyBox 3840x2160 (4 of 4096 spp) path d8", d5, 3840, 2160, 4096, 4)
    # VolPathIntegrator (the integrator Main/main.cpp:399 builds): same dragon-class mesh without media, as the boundary of a
    # HomogeneousMedium without material ("smoke"), and as a glass container of the medium (main.cpp:346-348)
    V = A.INTEGRATOR_VOLPATH
    run("6a dragon-class VolPath (no media) 1080p (16 of 1024 spp) d8", scenes.dragon_class(tess=467), 1920, 1080, 1024, 16, integrator=V)
    run("6b dragon-class smoke volume VolPath 1080p (16 of 1024 spp) d8", scenes.dragon_class(tess=467, volume="smoke"), 1920, 1080, 1024, 16, integrator=V)
    run("6c dragon-class glass + medium VolPath 1080p (16 of 1024 spp) d8", scenes.dragon_class(tess=467, volume="glass"), 1920, 1080, 1024, 16, integrator=V)
