#!/usr/bin/env python
"""Extract per-launch DRAM traffic / throughput / pipe utilisation from `.ncu-rep` files (ncu --set full) into JSON.
Usage: ncu_traffic.py OUT.json REP [REP ...]   (runs `ncu -i REP --page raw --csv` — works without a GPU)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__time_duration.sum": "duration", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct", "lts__t_sectors.sum": "l2_sectors",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__t_bytes.sum": "l2_bytes", "smsp__inst_executed.sum": "warp_insts",
    "sm__inst_executed_pipe_fma.sum": "fma_pipe_insts", "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__grid_size": "grid", "launch__block_size": "block",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
        "ns": 1e-3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}

out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        name = rec.get("Kernel Name", "?").split("(")[0].replace("void ", "")
        e = {}
        for k, short in WANT.items():
            if k in rec and rec[k] != "":
                try:
                    v = float(rec[k].replace(",", ""))
                except ValueError:
                    continue
                u = units[hdr.index(k)]
                e[short] = v * UNIT.get(u, 1.0)
        if "dram_read_bytes" in e and "dram_write_bytes" in e:
            e["dram_bytes"] = e["dram_read_bytes"] + e["dram_write_bytes"]
        out.setdefault(name, []).append(e)
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, v in out.items():
    for e in v:
        print(f"{k[:70]:70s} {e.get('duration', 0):8.1f} us  dram {e.get('dram_bytes', 0) / 1e6:8.1f} MB "
              f"({e.get('dram_pct', 0):4.1f}%)  tensor {e.get('tensor_pct', 0):4.1f}%  regs {int(e.get('regs', 0))}")
