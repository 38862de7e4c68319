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
for _ in range(2):
    _, _, st = sc.render(integrator=A.INTEGRATOR_VOLPATH, max_depth=8, s0=0, s1=16, out=out)
print("ms", st.ms_total, "launches", st.kernel_launches)
