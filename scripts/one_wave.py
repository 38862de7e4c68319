"""One 16-spp 1920x1080 wave of the bench workload (BASELINE configs[1]) - the program profiled under ncu.
Writes the ray counts of that wave to gpurun_out/one_wave.json so that scripts/ncu_counters.py can turn ncu's per-launch
counters into per-ray figures.  usage: one_wave.py [spp] [sampler: halton|sobol]"""
import json
import os
import subprocess
import sys

sys.path.insert(0, ".")
import numpy as np
from pysicalbasedraytracer_b200 import api, scenes

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H = 1920, 1080
sc = api.Scene(scenes.dragon_class(tess=467))
sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
_, _, st = sc.render(max_depth=8, rr_threshold=1.0, s0=0, s1=spp)
try:
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or None
except Exception:
    commit = None
out = dict(spp=spp, camera_samples=int(st.camera_samples), rays_closest=int(st.rays_closest), rays_shadow=int(st.rays_shadow),
           rays_path=int(st.rays_path), kernel_launches=int(st.kernel_launches), ms_total=float(st.ms_total), commit=commit)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/one_wave.json", "w") as f:
    json.dump(out, f)
print(json.dumps(out))
