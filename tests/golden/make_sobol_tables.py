"""Generates tests/golden/sobol_tables_1080p.npz: the reference's OWN Sobol' generator tables (Sampler/SobolMatrices.cpp:
SobolMatrices32, VdCSobolMatrices / VdCSobolMatricesInv for log2(resolution) = 11, i.e. sample bounds up to 2048 x 2048),
read out of the compiled reference (oracle/_ref).  They are DATA the caller of the C-ABI supplies through
pbrgpu_sampler_desc (the reference has the tables but no sampler that uses them; see DESIGN.md "Sobol: parity unpinned").
Only the first 256 dimensions are kept (a depth-8 path uses 5 + 8 * 8).  Run here, where /root/reference was compiled:
    python tests/golden/make_sobol_tables.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refharness as R  # noqa: E402

N_DIMS = 256
m32, vdc, vdi = R.sobol_tables((0, 0, 1920, 1080))
out = os.path.join(ROOT, "tests", "golden", "sobol_tables_1080p.npz")
np.savez_compressed(out, matrices32=m32[:N_DIMS * 52].copy(), vdc=vdc, vdc_inv=vdi, log2_resolution=np.int32(11))
print("wrote", out, os.path.getsize(out), "bytes")
