"""rays served by the bounce-1 k_trace launch of a 16-spp 1080p wave (the launch captured by ncu --set full)"""
import sys
sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes
data = scenes.dragon_class(tess=467)
sc = api.Scene(data)
sc.set_camera(api.CreatePerspectiveCamera(1920, 1080, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
sc.set_sampler(api.HaltonSampler(1024, (0, 0, 1920, 1080)))
_, _, st = sc.render(max_depth=1, rr_threshold=1.0, s0=0, s1=16)
print("launch1_rays", st.rays_closest + st.rays_shadow - st.camera_samples, "camera", st.camera_samples)
