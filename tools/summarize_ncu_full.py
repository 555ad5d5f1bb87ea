"""Selected metrics of an `ncu --set full` report -> markdown (profiles/<tag>_<name>.md)."""
import csv
import io
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
rep, tag, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor subpipe cycles (sum of 4)"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
idx = {h: i for i, h in enumerate(hdr)}
name_i = idx["Kernel Name"]
out = [f"# {title}", "", f"Source: `{Path(rep).name}` (`ncu --set full --clock-control none --import-source on`, one warm "
       "forward step of bench.py --profile-step; times under ncu are serialised/cold-cache).", ""]
out.append("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
out.append("|---|" + "---:|" * len(data))
out.append("| kernel | " + " | ".join(r[name_i].split("(")[0][-40:] for r in data) + " |")
for key, label in want:
    cands = [h for h in hdr if h == key or h.endswith("." + key) or h.endswith(key)]
    if not cands:
        continue
    i = idx[cands[0]]
    vals = []
    for r in data:
        try:
            vals.append(f"{float(r[i]):,.2f}")
        except ValueError:
            vals.append(r[i])
    out.append(f"| {label} [{units[i]}] | " + " | ".join(vals) + " |")
p = REPO / "profiles" / f"{tag}.md"
p.write_text("\n".join(out) + "\n")
print(p.read_text())
