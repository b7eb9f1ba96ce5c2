"""Compact summary of an .ncu-rep (first kernel result per kernel name): duration, tensor-pipe activity, DRAM/L2 traffic,
occupancy limits.   python tools/ncu_summary.py report.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
res = []
for r in rows[2:]:
    d = {}
    for k in WANT:
        if k in hdr:
            i = hdr.index(k)
            v = r[i]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            d[k] = {"value": v, "unit": units[i]} if k != "Kernel Name" else v
    res.append(d)
print(json.dumps(res, indent=1))
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
