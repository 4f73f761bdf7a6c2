#!/usr/bin/env python
"""Regenerate profiles/ from the two ncu passes of B200_PROFILING.md:

  ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_rN.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
  tools/gpu_call_ncu_full.sh   (ncu --set full --clock-control none --import-source on, two whole steps = 26 launches)

usage: tools/make_profiles.py <round> [launches.csv] [prof.ncu-rep]
writes profiles/rN_launches.csv, rN_launch_summary.md, rN_ncu_set_full_selected.csv, rN_ncu_digest.md, traffic.json
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SELECTED = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


def short(name):
    return name.split("(")[0][:60]


def launch_summary(rnd, path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = collections.OrderedDict()
    nrand = 0
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        # bench.py's secondary cfg2 measurement (absent with --no-secondary) starts with its own torch.randn inputs:
        # the summary covers the cfg3 workload only, i.e. the launches before the third normal-distribution kernel
        if "distribution_elementwise" in r["Kernel Name"]:
            nrand += 1
            if nrand > 2:
                break
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        k = short(r["Kernel Name"])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += v
    allus = sum(v[1] for v in tot.values())
    out = ["# Round %s launch list summary (ncu --metrics gpu__time_duration.sum --clock-control none, "
           "bench.py --steps 3 --warmup 3, workload cfg3:\n3 warm-up + 3 timed + 3 per-kernel-timing steps; the secondary "
           "cfg2 launches that follow in the raw CSV are excluded)" % rnd, "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.1f | %.3f |" % (k, n, us, us / allus))
    open(os.path.join(ROOT, "profiles", "r%s_launch_summary.md" % rnd), "w").write("\n".join(out) + "\n")
    with open(os.path.join(ROOT, "profiles", "r%s_launches.csv" % rnd), "w") as f:
        f.write("".join(rows))
    return tot


def full_selected(rnd, rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    keep = ["Kernel Name"] + [m for m in SELECTED if m in col]
    with open(os.path.join(ROOT, "profiles", "r%s_ncu_set_full_selected.csv" % rnd), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[col[k]] for k in keep])
        for r in data:
            w.writerow([r[col[k]] for k in keep])

    def tobytes(v, unit):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    acc = collections.defaultdict(list)
    for r in data:
        name = r[col["Kernel Name"]]
        key = ("fwd_recurrent" if "lstmp_fwd_" in name else "bwd_recurrent" if "lstmp_bwd_" in name
               else "gemm" if ("gemm_hl_kernel" in name or "gemm_tc_kernel" in name) else None)
        if key is None:
            continue
        acc[key].append(tobytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) +
                        tobytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]))
    traffic = {k: sum(v) / len(v) for k, v in acc.items()}
    traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches captured by "
                        "ncu --set full (profiles/r%s_ncu_set_full_selected.csv); bench workload cfg3, round %s"
                        % (rnd, rnd))
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    return traffic


def ncu_summary(rnd, rep):
    """Human-readable digest of the --set full capture: one block per kernel (first captured launch of each)."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        try:
            return float(r[col[m]].replace(",", ""))
        except Exception:
            return None

    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = ["# Round %s ncu --set full digest (tools/make_profiles.py; source: gpurun_out/prof_r%s.ncu-rep, command in"
           % (rnd, rnd), "# tools/make_profiles.py's header; per-launch values of the FIRST captured launch of each kernel / grid)", ""]
    # one step of the workload, launch by launch (the capture covers two whole steps: the first half)
    out += ["## All captured launches of one step (cfg3: layer 1 forward, layer 2 forward, layer 2 backward + its GEMM group, "
            "layer 1 backward + its GEMM group)", "",
            "| # | kernel | grid | duration us | tensor pipe % | DRAM read + write MB | L2 hit % |", "|---|---|---|---|---|---|---|"]
    for i, r in enumerate(data[:(len(data) + 1) // 2]):
        g = lambda m: val(r, m)
        out.append("| %d | `%s` | %d | %.1f | %.1f | %.1f | %.0f |" % (
            i, r[col["Kernel Name"]].split("(")[0], g("launch__grid_size"), g("gpu__time_duration.sum"),
            g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0.0,
            (g("dram__bytes_read.sum") or 0.0) + (g("dram__bytes_write.sum") or 0.0), g("lts__t_sector_hit_rate.pct") or 0.0))
    out.append("")
    seen = set()
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0]
        # the same kernel on another problem (input GEMM vs a layer's GEMM group) is told apart by grid and DRAM bytes
        key = (name, r[col["launch__grid_size"]], int(val(r, "dram__bytes_read.sum") or 0))
        if key in seen:
            continue
        seen.add(key)
        g = lambda m: val(r, m)
        out.append("## `%s` (grid %s, launch #%d of the table)" % (name, r[col["launch__grid_size"]], data.index(r)))
        out.append("")
        out.append("| metric | value |")
        out.append("|---|---|")
        dur = g("gpu__time_duration.sum")
        unit = units[col["gpu__time_duration.sum"]]
        out.append("| duration | %.1f %s |" % (dur, unit))
        out.append("| grid x block, regs/thread, dynamic smem | %d x %d, %d, %.1f KB |" % (
            g("launch__grid_size"), g("launch__block_size"), g("launch__registers_per_thread"),
            g("launch__shared_mem_per_block_dynamic")))
        rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
        out.append("| DRAM read / write | %.2f %s / %.2f %s (%.1f %% of peak throughput) |" % (
            rd, units[col["dram__bytes_read.sum"]], wr, units[col["dram__bytes_write.sum"]],
            g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
        out.append("| L2 hit rate | %.1f %% |" % g("lts__t_sector_hit_rate.pct"))
        out.append("| issue slots busy | %.1f %% |" % g("smsp__issue_active.avg.pct_of_peak_sustained_active"))
        out.append("| tensor pipe / FMA pipe active | %.1f %% / %.1f %% |" % (
            g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0.0,
            g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") or 0.0))
        wf, bc = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        if wf:
            out.append("| shared-memory wavefronts (LSU) / of which bank conflicts | %.3g / %.3g (%.0f %%) |" % (wf, bc, 100.0 * bc / wf))
        top = sorted(((g(m) or 0.0, m) for m in stalls), reverse=True)[:5]
        out.append("| top stall reasons (warps per issue) | %s |" % ", ".join(
            "%s %.2f" % (m[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for v, m in top))
        out.append("")
    open(os.path.join(ROOT, "profiles", "r%s_ncu_digest.md" % rnd), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "1"
    lcsv = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches_r%s.csv" % rnd)
    rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof_r%s.ncu-rep" % rnd)
    print(launch_summary(rnd, lcsv))
    print(full_selected(rnd, rep))
    ncu_summary(rnd, rep)
