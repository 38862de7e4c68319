"""One VolPath render of the dragon-class glass-with-medium scene (for ncu launch lists)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

W, H = 1920, 1080
sc = api.Scene(scenes.dragon_class(tess=467, volume=sys.argv[1] if len(sys.argv) > 1 else "glass"))
sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
out = np.zeros((H, W, 3), dtype=np.float32)
for flags in (0, 0, A.RENDER_TIMERS):
    _, _, st = sc.render(integrator=A.INTEGRATOR_VOLPATH, max_depth=8, s0=0, s1=16, out=out, flags=flags)
    print("ms %.2f launches %d | raygen %.2f closest %.2f shade %.2f mis %.2f film %.2f | Msamples/s %.1f" % (
        st.ms_total, st.kernel_launches, st.ms_raygen, st.ms_trace_closest, st.ms_shade, st.ms_trace_mis, st.ms_film,
        st.camera_samples / st.ms_total / 1e3), flush=True)
