"""Build a kernel-variant copy of the library (development aid): build_variant.py <tag> [-DNAME=VALUE ...]
-> build_variants/libpbrgpu_<tag>.so (run with PBRGPU_LIB=<that path>).  Objects of a variant live in build/obj_<tag>."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G

tag, defs = sys.argv[1], sys.argv[2:]
obj_dir = os.path.join(G.ROOT, "build", "obj_" + tag)
out_dir = os.path.join(G.ROOT, "build_variants")
os.makedirs(obj_dir, exist_ok=True)
os.makedirs(out_dir, exist_ok=True)
only = os.environ.get("VARIANT_ONLY")  # comma list of source name fragments that need the defines (others reuse build/obj)
jobs, objs = [], []
for src in G._sources():
    name = os.path.basename(src)
    if only and not any(f in name for f in only.split(",")):
        objs.append(os.path.join(G.OBJ, name + ".o"))
        continue
    obj = os.path.join(obj_dir, name + ".o")
    objs.append(obj)
    jobs.append((src, obj))


def comp(j):
    r = subprocess.run(["nvcc"] + G.NVCC_FLAGS + defs + ["-c", "-o", j[1], j[0]], cwd=G.CSRC, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)


with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
    list(ex.map(comp, jobs))
lib = os.path.join(out_dir, f"libpbrgpu_{tag}.so")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", lib] + objs + ["-Xlinker", "--no-as-needed", "-ldl"], check=True)
print("built", lib)
