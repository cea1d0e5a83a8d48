#!/usr/bin/env python
"""Condenses an .ncu-rep (read here with `ncu -i`, no GPU needed) into the text summaries kept under profiles/:
per-kernel headline metrics (raw page) and the hottest SASS regions with their average active threads (source page).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        # the L1 data pipe: the limiter of the traversal kernels (profiles/r01_l1_data_pipe.txt) — one wavefront per cycle and SM
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("# %s" % rep)
    name_i = hdr.index("Kernel Name")
    for k, r in enumerate(data):
        print("\n## launch %d: %s" % (k, r[name_i][:110]))
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print("  %-72s %s %s" % (key, r[i], units[i]))
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.2:
                    print("  stall %-66s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    src = page(rep, "source")
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and r and r[0].startswith("0x"):
            cur["rows"].append(r)
    for b in blocks[:1]:
        h = b["hdr"]
        iI, iT, iS, iSrc = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples"), h.index("Source")
        tot_i = sum(int(r[iI]) for r in b["rows"]) or 1
        tot_t = sum(int(r[iT]) for r in b["rows"])
        tot_s = sum(int(r[iS]) for r in b["rows"]) or 1
        print("\n## SASS regions of %s\n   %d instructions, %d warp-instructions executed, %.2f threads active per instruction on average" % (b["name"][:90], len(b["rows"]), tot_i, tot_t / tot_i))
        runs = []
        for k, r in enumerate(b["rows"]):
            ii, tt = int(r[iI]), int(r[iT])
            if runs and runs[-1]["ii"] == ii and runs[-1]["end"] == k - 1:
                runs[-1]["end"], runs[-1]["tt"], runs[-1]["n"], runs[-1]["s"] = k, runs[-1]["tt"] + tt, runs[-1]["n"] + 1, runs[-1]["s"] + int(r[iS])
            else:
                runs.append(dict(start=k, end=k, ii=ii, tt=tt, n=1, s=int(r[iS]), src=r[iSrc].strip()[:44]))
        print("   region(SASS idx)  n   share-of-warp-inst  avg-threads  stall-samples  first instruction")
        for R in sorted(runs, key=lambda R: -R["ii"] * R["n"])[:14]:
            print("   %4d-%4d      %3d        %5.2f%%          %5.1f        %5.2f%%     %s" % (R["start"], R["end"], R["n"], 100 * R["ii"] * R["n"] / tot_i, R["tt"] / max(1, R["ii"] * R["n"]), 100 * R["s"] / tot_s, R["src"]))


if __name__ == "__main__":
    main(sys.argv[1])
