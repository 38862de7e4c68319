"""k_trace tuning sweep (development aid): one 16-spp 1080p wave of the bench workload per setting of the environment knobs
the library reads when a scene renders for the first time.  usage: tune_trace.py "K1=v1,K2=v2" "K1=v3" ..."""
import os
import sys

sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

data = scenes.dragon_class(tess=467)
W, H = 1920, 1080
cam = api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0))))
for combo in (sys.argv[1:] or [""]):
    kv = dict(x.split("=") for x in combo.split(",") if x)
    for k, v in kv.items():
        os.environ[k] = v
    sc = api.Scene(data)
    sc.set_camera(cam)
    sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
    best = None
    for rep in range(3):
        _, _, st = sc.render(max_depth=8, rr_threshold=1.0, s0=16 * rep, s1=16 * rep + 16, flags=A.RENDER_TIMERS)
        if best is None or st.ms_total < best.ms_total:
            best = st
    _, _, sc_ = sc.render(max_depth=8, rr_threshold=1.0, s0=0, s1=16, flags=A.RENDER_COUNTERS)
    rays = sc_.rays_closest + sc_.rays_shadow
    print(f"{combo or 'default':40s} total {best.ms_total:6.2f} ms  closest {best.ms_trace_closest:6.2f}  shade {best.ms_shade:6.2f}  "
          f"nodes/ray {sc_.nodes_visited / rays:.2f} tris/ray {sc_.tris_tested / rays:.2f}", flush=True)
    sc.close()
    for k in kv:
        del os.environ[k]
