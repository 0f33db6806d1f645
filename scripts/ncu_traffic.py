"""Turn an `ncu --set full` capture of the dominant kernel into profiles/traffic.json (read by bench.py):

    ncu --set full --clock-control none --import-source on -k regex:bca_batch_dense_kernel --launch-skip 30 -c 1 \
        -o gpurun_out/r02_bca_batch_dense python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu --no-secondary
    python scripts/ncu_traffic.py gpurun_out/r02_bca_batch_dense.ncu-rep profiles/r02_bca_batch_dense_full.csv

Writes the raw page as CSV (committed under profiles/) and the per-launch DRAM traffic together with a hash of the
kernel's sources, so that bench.py can tell a capture of this build from a stale one."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

rep, out_csv = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
open(out_csv, "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def metric(name):
    v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return v * scale


rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
grid = vals[col["launch__grid_size"]]
info = {"kernel": vals[col["Kernel Name"]], "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
        "duration_us_under_ncu": metric("gpu__time_duration.sum") / 1e3 if units[col["gpu__time_duration.sum"]] == "ns"
        else float(vals[col["gpu__time_duration.sum"]]),
        "registers": int(float(vals[col["launch__registers_per_thread"]])), "grid": grid,
        "source_sha": kernel_source_sha(), "capture": os.path.basename(out_csv)}
json.dump({"bca_batch_dense_kernel": info["dram_bytes_per_launch"], **info}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"),
          indent=1)
print(json.dumps(info, indent=1))
