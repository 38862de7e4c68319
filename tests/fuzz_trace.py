"""One-off parity fuzz of the traversal kernels against the compiled reference (a checker, so it lives under tests/; not collected by
pytest; needs oracle/_ref and a GPU; run from the repository root: python tests/fuzz_trace.py 30):
N random closest-hit and any-hit rays on the full 872 k-triangle dragon-class mesh - uniform directions, near-axis-parallel
directions (either side of the fp16 / fp32 box-test switch), origins outside and inside the mesh - through k_trace's queues and
through the one-thread-per-ray helper loop.  usage: fuzz_trace.py [millions of rays]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes
from oracle import refharness as R  # test infrastructure (the compiled reference)

n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(2026)
data = scenes.dragon_class(tess=467)
sc = api.Scene(data)
rs = R.RefScene(data)
o = rng.uniform(-2.4, 2.4, (n, 3)).astype(np.float32)
o[n // 2:] *= 0.35
d = rng.normal(size=(n, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
k = n // 3  # a third of the rays: one (or two) direction components 2^-6 .. 2^-16 of the largest
small = rng.integers(0, 3, k)
ratio = np.exp2(-rng.uniform(6, 16, k)).astype(np.float32)
d[np.arange(k), small] = np.abs(d[:k]).max(axis=1) * ratio * rng.choice([-1.0, 1.0], k).astype(np.float32)
both = rng.random(k) < 0.3
idx = np.arange(k)[both]
d[idx, (small[both] + 1) % 3] = np.abs(d[idx]).max(axis=1) * ratio[both] * 0.5
rays = np.zeros((n, 8), dtype=np.float32)
rays[:, 0:3], rays[:, 3], rays[:, 4:7] = o, np.inf, d
t0 = time.time()
want = rs.trace(rays)
t1 = time.time()
print(f"reference: {n / (t1 - t0) / 1e6:.2f} Mrays/s, hit fraction {(want[:, 0] >= 0).mean():.3f}", flush=True)
for via in (True, False):
    got = sc.trace(rays, via_wave_kernel=via)
    bad = got[:, 0] != want[:, 0]
    same = ~bad
    tdiff = (got[same, 1].view(np.uint32) != want[same, 1].view(np.uint32)) & (want[same, 0] >= 0)
    print(("k_trace   " if via else "helper    ") + f"closest: id mismatches {bad.sum()} of {n} ({bad.mean():.2e}), t not bit-identical where ids agree: {tdiff.sum()}",
          flush=True)
    for i in np.nonzero(bad)[0][:20]:  # a tie (two triangles at the same t: shared edge) or a lost hit?
        print(f"    ray {i}: ids {int(got[i, 0])} / {int(want[i, 0])}, t {got[i, 1]!r} / {want[i, 1]!r}" +
              ("  (same t: a tie between two triangles)" if got[i, 1] == want[i, 1] else "  DIFFERENT t"), flush=True)
rays[:, 3] = rng.uniform(0.05, 3.0, n).astype(np.float32)
occ = rs.trace(rays, any_hit=True)
for via in (True, False):
    got = sc.trace(rays, any_hit=True, via_wave_kernel=via)
    print(("k_trace   " if via else "helper    ") + f"any-hit: mismatches {(got != occ).sum()} of {n} ({(got != occ).mean():.2e}), occluded fraction {occ.mean():.3f}", flush=True)
