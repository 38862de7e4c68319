"""Experiment: do two independent half-waves on two streams (trace of one overlapping shading of the other) beat one
wave?  Two Scene objects on one context, each rendering its own sample range into its own accumulator.
usage: overlap_probe.py <spp per scene> <n scenes>      env: PBRGPU_TRACE_BLOCKS_PER_SM"""
import sys
sys.path.insert(0, ".")
import torch
from pysicalbasedraytracer_b200 import api, scenes

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nsc = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, H = 1920, 1080
ctx = api.Context(0)
data = scenes.dragon_class(tess=467)
dev = torch.device("cuda", 0)
S = []
for k in range(nsc):
    sc = api.Scene(data, ctx)
    sc.set_camera(api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt((0, 0, 5), (0, 0, 0), (0, 1, 0)))))
    sc.set_sampler(api.HaltonSampler(1024, (0, 0, W, H)))
    S.append((sc, torch.zeros((H, W, 4), dtype=torch.float32, device=dev), torch.cuda.Stream(dev)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 6
for rep in range(3):
    torch.cuda.synchronize()
    e0.record(S[0][2])
    for k in range(1, nsc):
        S[k][2].wait_stream(S[0][2])
    for i in range(steps):
        for k, (sc, acc, st) in enumerate(S):
            s0 = (i * nsc + k) * spp
            sc.render_device(acc.data_ptr(), stream=st.cuda_stream, max_depth=8, rr_threshold=1.0, s0=s0, s1=s0 + spp, want_stats=False)
    for k in range(1, nsc):
        S[0][2].wait_stream(S[k][2])
    e1.record(S[0][2])
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{nsc} scene(s) x {spp} spp per call: {ms / steps:.2f} ms per {nsc * spp} spp  ->  {W * H * spp * nsc * steps / ms / 1e3:.1f} Msamples/s", flush=True)
