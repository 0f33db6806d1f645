"""Per-source-line summary of an `ncu --set full --import-source on` capture: share of warp-stall samples and of
executed warp instructions per CUDA source line (innermost inlined frame), as CSV on stdout.

    python scripts/ncu_source_lines.py gpurun_out/r02_targets.ncu-rep bca_exact_dense_cluster_kernel [label] > out.csv
"""
import collections
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
label = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kernel], capture_output=True, text=True, check=True).stdout


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


hdr, cur_file, cur_line, seen = None, "", None, {}
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr is not None and len(r) >= 8:
        if r[0] != "":
            cur_line = (cur_file, r[0], " ".join(r[1].split())[:100])
        elif r[2] not in seen:          # one SASS instruction appears under every inlined frame: count it once
            seen[r[2]] = (cur_line, num(r[6]), num(r[7]), r[3].strip())
tot_s = sum(v[1] for v in seen.values()) or 1.0
tot_i = sum(v[2] for v in seen.values()) or 1.0
samples, instr = collections.Counter(), collections.Counter()
for line, s, i, _ in seen.values():
    samples[line] += s
    instr[line] += i
w = csv.writer(sys.stdout)
w.writerow(["capture", "pct_of_stall_samples", "pct_of_warp_instructions", "file", "line", "source"])
for line, s in samples.most_common(40):
    w.writerow([label, f"{100 * s / tot_s:.2f}", f"{100 * instr[line] / tot_i:.2f}", line[0], line[1], line[2]])
w.writerow([label, "total_samples", int(tot_s), "total_warp_instructions", int(tot_i), ""])
top = sorted(seen.values(), key=lambda v: -v[1])[:12]
for line, s, i, sass in top:
    w.writerow([label + " sass", f"{100 * s / tot_s:.2f}", "", line[0], line[1], sass[:80]])
