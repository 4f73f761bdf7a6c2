#!/usr/bin/env python
"""Source-line hot spots of the --set full capture: warp-stall samples per CUDA source line (file:line) of the first
captured launch of each kernel, with the stall reasons of that line -> profiles/r<N>_ncu_source_hotspots.md.

usage: python tools/ncu_hotspots.py <round> [prof.ncu-rep]        (needs the library built with -lineinfo and the
capture taken with --import-source on; reads the report here, no GPU)
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = [("lstmp_fwd_tma_kernel", 0), ("lstmp_bwd_tma_kernel", 0), ("gemm_hl_kernel", 2), ("split_multi_kernel", 0)]
TOP = 14


def page(rep, kernel, skip):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kernel, "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    files = []  # (path, header, rows)
    cur = None
    for row in csv.reader(io.StringIO(txt)):
        if len(row) == 2 and row[0] == "File Path":
            cur = [row[1], None, []]
            files.append(cur)
        elif cur is not None and row and row[0] == "Line No":
            cur[1] = row
        elif cur is not None and cur[1] is not None and len(row) == len(cur[1]) and row[0].strip().isdigit():
            cur[2].append(row)   # a source line with the metrics of its SASS aggregated (SASS rows have no line number)
    return files


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "2"
    rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "prof_r%s.ncu-rep" % rnd)
    out = ["# Round %s: warp-stall samples per source line (ncu --set full --import-source on; tools/ncu_hotspots.py)" % rnd, "",
           "One block per kernel (one captured launch each; for `gemm_hl_kernel` the layer-2 backward group). `samples` = warp-stall",
           "sampling hits on the SASS of that line (all samples); `share` = of the kernel's samples; `stalls` = the reasons of",
           "those samples. A time loop's samples sit where its warps WAIT: the mbarrier waits of the pipeline, the CTA",
           "barriers around them and the single polling thread of the grid barrier.", ""]
    for kernel, skip in KERNELS:
        files = page(rep, kernel, skip)
        lines = []
        total = 0
        for path, hdr, rows in files:
            col = {h: i for i, h in enumerate(hdr)}
            isamp = col.get("# Samples")
            stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
            for r in rows:
                try:
                    n = int(r[isamp].replace(",", ""))
                except Exception:  # noqa: BLE001
                    continue
                total += n
                if n == 0:
                    continue
                st = collections.Counter()
                for h, i in stall_cols:
                    try:
                        v = int(r[i].replace(",", ""))
                    except Exception:  # noqa: BLE001
                        v = 0
                    if v:
                        st[h[len("stall_"):]] = v
                lines.append((n, os.path.relpath(path, ROOT) if path.startswith(ROOT) else path, r[0], r[1].strip(), st))
        lines.sort(key=lambda x: -x[0])
        out.append("## `%s` (%d samples)" % (kernel, total))
        out.append("")
        out.append("| samples | share | file:line | source | stalls |")
        out.append("|---|---|---|---|---|")
        for n, path, ln, src, st in lines[:TOP]:
            out.append("| %d | %.1f %% | `%s:%s` | `%s` | %s |" % (
                n, 100.0 * n / max(total, 1), path.replace("kaldi-lstm_b200/csrc/", ""), ln, src[:90].replace("|", "\\|"),
                ", ".join("%s %d" % kv for kv in st.most_common(3))))
        out.append("")
    dst = os.path.join(ROOT, "profiles", "r%s_ncu_source_hotspots.md" % rnd)
    open(dst, "w").write("\n".join(out) + "\n")
    print(dst)


if __name__ == "__main__":
    main()
