#!/usr/bin/env python
"""bench.py - headline benchmark of the GPU PathIntegrator hot path (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Stanford-Dragon-class synthetic mesh (872 368 triangles) inside the
Cornell walls, Matte + Mirror, Halton sampler, 1920x1080, PathIntegrator maxDepth 8.  The config is quoted
at 1024 spp; a *step* is one pass of the hot path over one batch of that job: SPP_PER_STEP samples of every
pixel (sample indices advance from step to step, so steps are consecutive slices of the 1024-spp job).
Metric: Msamples/s (camera samples = PathIntegrator::Li calls), with Mrays/s beside it.

  value  - device-resident: the accumulator stays in HBM (pbrgpu_render_device), CUDA events on the
           launching stream, barrier + synchronize on both sides, max over ranks.
  e2e    - the same step through the host-buffer C-ABI call pbrgpu_render (what Integrator::Render does):
           descriptors host->device, radiance image device->host every step, wall-clock around the call.
  N > 1  - scene replicated per GPU, sample ranges sharded across ranks (weak scaling: every rank renders
           SPP_PER_STEP samples per pixel per step), per-step NCCL reduce(sum) of the fp32 accumulators to rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
SPP_TOTAL = 1024
MAX_DEPTH = 8
RR_THRESHOLD = 1.0
TESS = 467
SPP_PER_STEP = 16
EYE, LOOK, UP = (0.0, 0.0, 5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)
WORKLOAD = "dragon-class 872368-tri mesh in Cornell walls, Matte+Mirror, Halton, 1920x1080, PathIntegrator maxDepth 8 (configs[1])"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for k, nme in enumerate(names):
                    if p[5 + k].lower().startswith("active"):
                        reasons.add(nme)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def build_scene_data():
    from pysicalbasedraytracer_b200 import scenes
    return scenes.dragon_class(tess=TESS)


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU PathIntegrator (oracle/_ref), all host threads
# --------------------------------------------------------------------------------------------------
def cpu_reference_steps(data, steps, warmup, budget_s):
    """Time the compiled reference (oracle/_ref) on a bounded sample of the same workload: full-width
    row bands of the 1920x1080 image at 1 spp, sized from a probe so one step takes ~budget_s/steps."""
    from oracle import refharness as R
    from pysicalbasedraytracer_b200 import _abi as A
    if not R.available():
        return None
    rs = R.RefScene(data)
    rc = R.RefCamera(A.CAMERA_PERSPECTIVE, W, H, EYE, LOOK, UP)
    cores = R.lib().ref_num_procs()
    kw = dict(integrator=A.INTEGRATOR_PATH, max_depth=MAX_DEPTH, rr_threshold=RR_THRESHOLD, threads=0)
    # probe: 8 rows through the middle of the mesh
    y0 = H // 2 - 4
    _, st = R.render(rs, rc, SPP_TOTAL, (0, 0, W, H), region=(0, y0, W, y0 + 8), s0=0, s1=1, **kw)
    rate = st.samples / max(st.seconds, 1e-9)
    per_step = max(budget_s / max(steps + warmup, 1), 0.5)
    rows = int(min(H, max(8, rate * per_step / W)))
    rows -= rows % 2
    ya = (H - rows) // 2
    times, samples, rays = [], 0, 0
    for i in range(warmup + steps):
        _, st = R.render(rs, rc, SPP_TOTAL, (0, 0, W, H), region=(0, ya, W, ya + rows), s0=i, s1=i + 1, **kw)
        if i >= warmup:
            times.append(st.seconds)
            samples += st.samples
            rays += st.rays_closest + st.rays_shadow
    total = sum(times)
    return dict(value=samples / total / 1e6, unit="Msamples/s", mrays_per_s=rays / total / 1e6, cores=int(cores),
                kind="reference",
                sample=f"{rows} full-width rows (y {ya}..{ya + rows}) of the 1920x1080 frame at 1 spp per step, "
                       f"{steps} steps, reference PathIntegrator::Li via oracle/_ref with {cores} OpenMP threads "
                       f"(dynamic row schedule); BVH build {rs.bvh_seconds():.2f} s excluded",
                ms_per_step=1e3 * total / max(len(times), 1), bvh_build_s=rs.bvh_seconds())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data = build_scene_data()
    r = cpu_reference_steps(data, args.steps, args.warmup, budget_s=90.0)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpbr_ref.so missing"}))
        return
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": r["value"], "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sampler": "Halton", "max_depth": MAX_DEPTH, "resolution": [W, H]},
        "mrays_per_s": r["mrays_per_s"],
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pysicalbasedraytracer_b200 import api
    from pysicalbasedraytracer_b200 import _abi as A

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if api.device_count() == 0:
        raise RuntimeError("bench.py: no CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    data = build_scene_data()
    ctx = api.Context(local)
    scene = api.Scene(data, ctx)
    cam = api.CreatePerspectiveCamera(W, H, api.Inverse(api.LookAt(EYE, LOOK, UP)))
    scene.set_camera(cam)
    scene.set_sampler(api.HaltonSampler(SPP_TOTAL, (0, 0, W, H)))
    kw = dict(integrator=A.INTEGRATOR_PATH, max_depth=MAX_DEPTH, rr_threshold=RR_THRESHOLD)

    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    # one explicit stream carries the render launches, the NCCL reduce and the timing events
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def sample_range(step):
        # consecutive slices of the 1024-spp job; ranks take disjoint sub-ranges of each slice
        base = (step * world + rank) * SPP_PER_STEP
        base %= max(SPP_TOTAL - SPP_PER_STEP, 1)
        return base, base + SPP_PER_STEP

    def step_device(step, flags=0, want_stats=False):
        s0, s1 = sample_range(step)
        st = scene.render_device(accum.data_ptr(), stream=stream.cuda_stream, s0=s0, s1=s1, flags=flags,
                                 want_stats=want_stats, **kw)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        return st

    # ---- warm-up ----
    for i in range(args.warmup):
        step_device(i)
    barrier()

    # ---- timed region: exactly K steps, device-resident ----
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    samples_per_step = W * H * SPP_PER_STEP * world
    value = samples_per_step * args.steps / (ms_max * 1e-3) / 1e6

    # ---- per-kernel timeline + counters (separate passes, same steps; not part of `value`) ----
    barrier()
    tsum = dict(closest=0.0, shade=0.0, shadow=0.0, mis=0.0, raygen=0.0, film=0.0, total=0.0, n_closest=0,
                rays_closest=0, rays_shadow=0, rays_path=0)
    for i in range(args.steps):
        st = step_device(args.warmup + i, flags=A.RENDER_TIMERS, want_stats=True)
        tsum["closest"] += st.ms_trace_closest; tsum["shade"] += st.ms_shade; tsum["shadow"] += st.ms_trace_shadow
        tsum["mis"] += st.ms_trace_mis; tsum["raygen"] += st.ms_raygen; tsum["film"] += st.ms_film
        tsum["total"] += st.ms_total; tsum["n_closest"] += st.n_trace_closest_launches
        tsum["rays_closest"] += st.rays_closest; tsum["rays_shadow"] += st.rays_shadow; tsum["rays_path"] += st.rays_path
    barrier()
    stc = step_device(args.warmup, flags=A.RENDER_COUNTERS, want_stats=True)
    barrier()
    rays_c = stc.rays_closest + stc.rays_shadow
    nodes_per_ray = stc.nodes_visited / max(rays_c, 1)
    tris_per_ray = stc.tris_tested / max(rays_c, 1)

    # ---- e2e: host-buffer C-ABI call, wall clock, H2D descriptors + D2H image inside the timed region ----
    barrier()
    # the application's frame buffer: allocated once, reused; page-locked so that the image leaves the GPU in one DMA
    host_film = torch.zeros((H, W, 3), dtype=torch.float32, pin_memory=True).numpy()
    for i in range(2):
        s0, s1 = sample_range(i)
        scene.render(s0=s0, s1=s1, out=host_film, **kw)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        s0, s1 = sample_range(args.warmup + i)
        rgb, _, _ = scene.render(s0=s0, s1=s1, out=host_film, **kw)
    torch.cuda.synchronize(dev)
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / float(te.item()) / 1e6
    import ctypes as C
    h2d = C.sizeof(A.RenderDesc) + C.sizeof(A.Stats)
    d2h = W * H * 3 * 4 + C.sizeof(A.Stats)

    if rank == 0:
        info = scene.info()
        # roofline of the dominant kernel (k_trace: path + shadow + MIS rays of a bounce): algorithmic bytes per ray
        #   B_ray = 32 (ray record) + 16 (hit record) + nodes/ray * S_node + tris/ray * S_tri   
        S_NODE, S_TRI = 80, 48
        b_ray = 32 + 16 + nodes_per_ray * S_NODE + tris_per_ray * S_TRI
        rays_per_launch = (tsum["rays_closest"] + tsum["rays_shadow"]) / max(tsum["n_closest"], 1)
        avg_ms = tsum["closest"] / max(tsum["n_closest"], 1)
        achieved = rays_per_launch * b_ray / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        peak, peak_src = _peaks()
        # DRAM traffic of the same kernel from the committed ncu --set full capture (per ray, scaled to the mean launch)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_k_trace_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = float(json.load(f)["dram_bytes_per_ray"]) * rays_per_launch
        rays_total = tsum["rays_closest"] + tsum["rays_shadow"]
        launches_per_step = int(stc.kernel_launches)
        cpu = cpu_reference_steps(data, 4, 1, budget_s=25.0) if world == 1 and not args.no_cpu else None
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sampler": "Halton", "max_depth": MAX_DEPTH, "resolution": [W, H],
                       "spp_per_step_per_gpu": SPP_PER_STEP, "spp_job": SPP_TOTAL, "sharding": f"sample-range x{world}",
                       "l2_policy": "working set per step (wave state 7.5 GB + 54 MB scene) exceeds the 126 MB L2; no explicit flush"},
            "mrays_per_s": rays_total / max(tsum["total"], 1e-9) / 1e3 * world,
            "rays_per_sample": rays_total / (W * H * SPP_PER_STEP * args.steps),
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_ray": b_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "rays_per_launch": rays_per_launch, "avg_launch_ms": avg_ms,
                         "note": "k_trace serves path + shadow + MIS rays of one bounce in one launch; the scene (12 MB compressed 8-wide nodes + 42 MB triangles) is L2-resident, so the algorithmic node/triangle bytes are served by L1/L2 (measured DRAM traffic is ~95 B/ray, profiles/r01_k_trace_traffic.json); ncu shows the kernel issue-slot bound (0.73 issue/cycle/SMSP, 17 of 32 lanes per instruction), not bandwidth bound"},
            "kernel_ms_per_step": {k: tsum[k] / args.steps for k in ("raygen", "closest", "shade", "mis", "film", "total")},
            "bvh": info,
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["mrays_per_s"] = cpu["mrays_per_s"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "gpu" and os.environ.get("PBRGPU_ALLOW_SHORT_WARMUP") != "1":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
