"""Turn an ncu metrics pass over scripts/one_wave.py into the per-ray counters bench.py reads (profiles/r02_k_trace_counters.json).

  on the GPU box:   python scripts/one_wave.py && ncu --metrics <METRICS> --clock-control none -k regex:k_trace --csv \
                        --log-file gpurun_out/ncu_counters.csv python scripts/one_wave.py
  here:             python scripts/ncu_counters.py gpurun_out/ncu_counters.csv gpurun_out/one_wave.json <commit> \
                        > profiles/r02_k_trace_counters.json

Sums every k_trace launch of the wave (one per bounce) and divides by the rays those launches served
(rays_closest + rays_shadow of the render).  Nothing here is typed in by hand."""
import csv
import json
import sys

METRICS = ("smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "gpu__time_duration.sum,lts__t_sectors_lookup_hit.sum,lts__t_sectors_lookup_miss.sum,lts__t_bytes.sum,"
           "l1tex__t_sectors_lookup_hit.sum,l1tex__t_sectors_lookup_miss.sum,smsp__cycles_active.sum,smsp__issue_active.sum,"
           "sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum")


def main():
    if len(sys.argv) < 3:
        print(METRICS)
        return
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iID = hdr.index("ID")
    pattern = sys.argv[4] if len(sys.argv) > 4 else "k_trace"
    tot, launches = {}, set()
    for r in rows[1:]:
        if len(r) <= iV or pattern not in r[iK]:
            continue
        launches.add(r[iID])
        v = float(r[iV].replace(",", ""))
        unit = r[iU].lower()
        if r[iM].startswith("gpu__time_duration"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        elif "byte" in unit:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        tot[r[iM]] = tot.get(r[iM], 0.0) + v
    wave = json.load(open(sys.argv[2]))
    rays = wave["rays_closest"] + wave["rays_shadow"]
    g = lambda k: tot.get(k, 0.0)
    hit2, miss2 = g("lts__t_sectors_lookup_hit.sum"), g("lts__t_sectors_lookup_miss.sum")
    hit1, miss1 = g("l1tex__t_sectors_lookup_hit.sum"), g("l1tex__t_sectors_lookup_miss.sum")
    out = {
        "kernel": pattern, "commit": sys.argv[3], "source": "ncu --metrics (scripts/ncu_counters.py METRICS) --clock-control none over scripts/one_wave.py",
        "launches": len(launches), "rays": rays, "camera_samples": wave["camera_samples"],
        "gpu_time_ms_under_ncu": g("gpu__time_duration.sum"),
        "warp_inst": g("smsp__inst_executed.sum"), "thread_inst": g("smsp__thread_inst_executed.sum"),
        "warp_inst_per_ray": g("smsp__inst_executed.sum") / rays, "thread_inst_per_ray": g("smsp__thread_inst_executed.sum") / rays,
        "lanes_per_inst": g("smsp__thread_inst_executed.sum") / max(g("smsp__inst_executed.sum"), 1),
        # warp instructions per ray by execution pipe (the ALU pipe takes one warp instruction per TWO cycles: scripts/ubench/pipes.cu)
        "alu_warp_inst_per_ray": g("sm__inst_executed_pipe_alu.sum") / rays, "fma_warp_inst_per_ray": g("sm__inst_executed_pipe_fma.sum") / rays,
        "lsu_warp_inst_per_ray": g("sm__inst_executed_pipe_lsu.sum") / rays, "xu_warp_inst_per_ray": g("sm__inst_executed_pipe_xu.sum") / rays,
        "issue_active_per_cycle_under_ncu": g("smsp__issue_active.sum") / max(g("smsp__cycles_active.sum"), 1),
        "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
        "dram_bytes_per_ray": (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / rays,
        "l2_bytes_per_ray": g("lts__t_bytes.sum") / rays,
        "l2_sector_hit_rate": hit2 / max(hit2 + miss2, 1), "l1_sector_hit_rate": hit1 / max(hit1 + miss1, 1),
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
