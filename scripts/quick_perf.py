"""Quick throughput probe (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A

tess = int(sys.argv[1]) if len(sys.argv) > 1 else 467
W, H, spp = 1920, 1080, int(sys.argv[2]) if len(sys.argv) > 2 else 8
t0 = time.time()
data = scenes.dragon_class(tess=tess)
print("scene gen", time.time() - t0, "tris", data.n_tris)
sc = api.Scene(data)
print("scene create", sc.build_seconds, sc.info())
sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
for flags in (0, 0, A.RENDER_COUNTERS):
    out, _, st = sc.render(max_depth=8, rr_threshold=1.0, s0=0, s1=spp, flags=flags)
    rays = st.rays_closest + st.rays_shadow
    print(f"ms {st.ms_total:.1f}  Msamples/s {st.camera_samples / st.ms_total / 1e3:.1f}  Mrays/s {rays / st.ms_total / 1e3:.1f} "
          f"rays/sample {rays / st.camera_samples:.2f} launches {st.kernel_launches} nodes/ray {st.nodes_visited / max(rays,1):.1f} tris/ray {st.tris_tested / max(rays,1):.1f} mean {out.mean():.4f}")
