"""Per-kernel-group times of the bench workload (16 spp of the dragon-class 1080p frame) - development aid."""
import sys

import numpy as np

sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

W, H, SPP = 1920, 1080, 16
S0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sc = api.Scene(scenes.dragon_class(tess=467))
sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
out = np.zeros((H, W, 3), dtype=np.float32)
for flags in (0, 0, A.RENDER_TIMERS, A.RENDER_TIMERS):
    _, _, st = sc.render(max_depth=8, s0=S0, s1=S0 + SPP, out=out, flags=flags)
    print("flags %d ms_total %.2f | raygen %.2f closest %.2f shade %.2f mis %.2f film %.2f | Msamples/s %.1f" %
          (flags, st.ms_total, st.ms_raygen, st.ms_trace_closest, st.ms_shade, st.ms_trace_mis, st.ms_film,
           st.camera_samples / st.ms_total / 1e3), flush=True)
