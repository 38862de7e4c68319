import sys
import numpy as np
sys.path.insert(0, ".")
from pysicalbasedraytracer_b200 import api, scenes, _abi as A
from oracle import refharness as R
from tests.util import Pair, rel_err
p = Pair(api, R, scenes.cornell(sphere_tess=16), 96, 96, 4, sampler_type=A.SAMPLER_SOBOL)
for depth in (0, 1):
    for s0 in range(4):
        want, rst = R.render(p.rscene, p.rcam, 4, (0, 0, 96, 96), max_depth=depth, rr_threshold=0.8, s0=s0, s1=s0 + 1, sampler_type=A.SAMPLER_SOBOL)
        got, _, st = p.scene.render(max_depth=depth, rr_threshold=0.8, s0=s0, s1=s0 + 1)
        err = rel_err(got, want)
        bad = np.argwhere(err > 1e-3)
        print("depth", depth, "sample", s0, "bad", len(bad), "rays", st.rays_closest, st.rays_shadow, rst.rays_closest, rst.rays_shadow)
        for (y, x) in bad[:6]:
            d = p.scene.dump_sampler((x, y, x + 1, y + 1), s0, s0 + 1, 12)[0, 0]
            dr = p.rsmp.dump((x, y, x + 1, y + 1), s0, s0 + 1, 12)[0, 0]
            print("  pix", x, y, "gpu", got[y, x], "ref", want[y, x], "dims", d[:12], "same", np.array_equal(d, dr))
