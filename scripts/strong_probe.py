"""Development aid: wall time of the fixed 1024-spp job (one pbrgpu_render call) on a group context of N devices."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
W, H = 1920, 1080
ctx = api.Context(devices=list(range(n))) if n > 1 else api.Context(0)
sc = api.Scene(scenes.dragon_class(tess=467), ctx)
sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
out = np.zeros((H, W, 3), dtype=np.float32)
for rep in range(3):
    for s1, rgba in ((64, False), (spp, False), (spp, True)):
        t0 = time.perf_counter()
        _, _, st = sc.render(max_depth=8, rr_threshold=1.0, s0=0, s1=s1, out=out, want_rgba8=rgba)
        dt = time.perf_counter() - t0
        print(f"N={n} spp={s1} rgba8={rgba}: wall {dt*1e3:.1f} ms, device ms_total {st.ms_total:.1f}, {W*H*s1/dt/1e6:.1f} Msamples/s", flush=True)
