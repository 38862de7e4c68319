"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/pbrgpu.h declares, its
POD structs match the ctypes mirror byte for byte, the host-only helpers agree with the oracle, and the
compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from pysicalbasedraytracer_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pbrgpu.h")


def test_library_exports_every_declared_symbol(api):
    lib = api.lib()
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(pbrgpu_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libpbrgpu.so does not export {n}"


def test_struct_layouts_match_header():
    structs = {"pbrgpu_material": A.Material, "pbrgpu_light": A.Light, "pbrgpu_medium": A.Medium, "pbrgpu_scene_desc": A.SceneDesc,
               "pbrgpu_camera_desc": A.CameraDesc, "pbrgpu_sampler_desc": A.SamplerDesc,
               "pbrgpu_render_desc": A.RenderDesc, "pbrgpu_stats": A.Stats, "pbrgpu_dump_desc": A.DumpDesc}
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "pbrgpu.h"\nint main(){\n'
    for name, cls in structs.items():
        src += f'printf("{name} %zu\\n", sizeof({name}));\n'
        for fname, _ in cls._fields_:
            src += f'printf("{name}.{fname} %zu\\n", offsetof({name}, {fname}));\n'
    src += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for fname, _ in cls._fields_:
            assert int(got[f"{name}.{fname}"]) == getattr(cls, fname).offset, f"{name}.{fname}"


def test_host_helpers_match_oracle(api):
    from oracle import restate as RS
    for b in [(0, 0, 512, 512), (0, 0, 1920, 1080), (0, 0, 20, 9), (0, 0, 3840, 2160), (5, 7, 100, 60), (0, 0, 1, 1)]:
        assert api.halton_params(b) == RS.halton_params(b)
    assert np.array_equal(api.compute_perms(), RS.perms())


def test_no_cpu_fallback(api):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.PbrGpuError, match="no CUDA device"):
        api.Context(0)
    from pysicalbasedraytracer_b200 import scenes
    with pytest.raises(api.PbrGpuError):
        api.Scene(scenes.cornell(sphere_tess=0))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or execute it."""
    pkg = os.path.join(ROOT, "pysicalbasedraytracer_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|refharness|restate|libpbr_ref|librestate", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(text), f"{f} references the oracle: the product path must never route through it"
