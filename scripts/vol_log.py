"""Per-iteration VolPath queue sizes / k_trace times (PBRGPU_VOL_LOG=1) on the dragon-class smoke volume."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
os.environ["PBRGPU_VOL_LOG"] = "1"
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

W, H, SPP = 1920, 1080, 16
data = scenes.dragon_class(tess=467, volume=sys.argv[1] if len(sys.argv) > 1 else "smoke")
sc = api.Scene(data)
sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
out = np.zeros((H, W, 3), dtype=np.float32)
_, _, st = sc.render(integrator=A.INTEGRATOR_VOLPATH, max_depth=8, s0=0, s1=SPP, out=out, flags=0 if (len(sys.argv) > 2 and sys.argv[2] == "plain") else A.RENDER_COUNTERS)
print("ms_total", st.ms_total)
