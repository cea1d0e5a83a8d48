#!/usr/bin/env python
"""Reads .ncu-rep files (here, with `ncu -i`; no GPU needed) and writes the counters bench.py's roofline blocks cite into
profiles/r02_counters.json: per kernel, what names its limiter — L1 data-pipe wavefronts, issue slots, L2 and DRAM traffic.
usage: python tools/ncu_counters.py gpurun_out/r02_ncu_raycast.ncu-rep gpurun_out/r02_ncu_pt.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {"duration_ms": "gpu__time_duration.sum", "registers": "launch__registers_per_thread", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
     "warp_inst": "smsp__inst_executed.sum", "l1_lsu_wavefront_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
     "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct", "l2_bytes": "lts__t_bytes.sum",
     "l2_throughput_pct": "lts__t_bytes.avg.pct_of_peak_sustained_elapsed", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
     "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
     "global_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global_ld_requests": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def record(h, u, r):
    d = {"kernel": r[h.index("Kernel Name")][:120]}
    for k, m in M.items():
        if m not in h:
            continue
        try:
            v = float(r[h.index(m)].replace(",", ""))
        except ValueError:
            continue
        d[k] = v * SCALE.get(u[h.index(m)], 1.0)
    if "dram_read" in d and "dram_write" in d:
        d["dram_bytes"] = d["dram_read"] + d["dram_write"]
    return d


def main(reps):
    out = {"source": "ncu --set full --clock-control none --import-source on (tools/r02_measure.sh); per-launch values are cold-cache and serialised"}
    for rep in reps:
        h, u, data = rows(rep)
        recs = [record(h, u, r) for r in data]
        if "raycast" in rep:  # rb_k_trace_phased<ANY, ORDERED, BLOCKS>: closest = <0, 0, 0>, any = <1, 1, 0>; of each kind the longer launch is the incoherent set
            import re
            for kind, first in (("closest", "0"), ("any", "1")):  # the first template argument is ANY, printed as `0` or `(bool)0`
                c = sorted([r for r in recs if re.search(r"rb_k_trace_phased<(?:\(bool\))?" + first + r"\b", r["kernel"])], key=lambda r: r.get("duration_ms", 0))
                if len(c) >= 2:
                    out["raycast_%s_coherent" % kind], out["raycast_%s_incoherent" % kind] = c[0], c[-1]
        else:  # path tracer: the longest launch of each kernel class (the first bounce of a full batch)
            for cls in ("trace", "shade", "shadow"):
                c = [r for r in recs if "rb_k_pt_" + cls in r["kernel"]]
                if c:
                    out["pathtrace_" + cls] = max(c, key=lambda r: r.get("duration_ms", 0))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_counters.json"), "w"), indent=1)
    print(json.dumps({k: (v if isinstance(v, str) else {a: v[a] for a in ("duration_ms", "l1_lsu_wavefront_pct", "issue_active_pct", "dram_bytes", "l2_hit_pct") if a in v}) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
