"""VolPath per-kernel-group times and iteration counts on the dragon-class volume scenes (development aid)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

W, H, SPP = 1920, 1080, 16
for vol in (None, "smoke", "glass"):
    data = scenes.dragon_class(tess=467, volume=vol)
    sc = api.Scene(data)
    sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
    sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
    out = np.zeros((H, W, 3), dtype=np.float32)
    for flags in (0, A.RENDER_TIMERS, A.RENDER_COUNTERS):
        for _ in range(2):
            _, _, st = sc.render(integrator=A.INTEGRATOR_VOLPATH, max_depth=8, s0=0, s1=SPP, out=out, flags=flags)
        print(vol, "flags", flags, "ms_total %.2f launches %d trace-launches %d | raygen %.2f closest %.2f shade %.2f mis %.2f film %.2f | rays %d path %d nodes/ray %.2f tris/ray %.2f"
              % (st.ms_total, st.kernel_launches, st.n_trace_closest_launches, st.ms_raygen, st.ms_trace_closest, st.ms_shade,
                 st.ms_trace_mis, st.ms_film, st.rays_closest, st.rays_path, st.nodes_visited / max(1, st.rays_closest),
                 st.tris_tested / max(1, st.rays_closest)), flush=True)
    sc.close()
