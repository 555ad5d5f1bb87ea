"""Turns the ncu CSV of one profiled bench step (gpurun_out/launches.csv: gpu__time_duration.sum +
dram bytes per launch) into profiles/<tag>_launches.md and profiles/gemm_traffic.json."""
import collections
import csv
import json
import re
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
src = Path(sys.argv[1] if len(sys.argv) > 1 else REPO / "gpurun_out" / "launches.csv")
tag = sys.argv[2] if len(sys.argv) > 2 else "r01"
lines = [l for l in src.read_text().splitlines() if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    key = r["ID"]
    d = per.setdefault(key, {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("mts::", "")})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    elif m.startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[m] = v * mult
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    a = agg[d["name"]]
    a[0] += 1
    a[1] += d.get("us", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
total = sum(a[1] for a in agg.values())
out = [f"# ncu launch list of one profiled forward step ({tag})", "",
       "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
       "--profile-from-start off python bench.py --profile-step` (bidmc_llama2_7b, one warm step). Per-launch times "
       "under ncu are serialised and cold-cache: compare SHARES, not absolutes.", "",
       "| kernel | launches | total ms | share | DRAM GB (r+w) |", "|---|---:|---:|---:|---:|"]
for name, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{name}` | {n} | {us / 1e3:.3f} | {100 * us / total:.1f} % | {by / 1e9:.2f} |")
out.append(f"| **total** | {sum(a[0] for a in agg.values())} | {total / 1e3:.3f} | 100 % | {sum(a[2] for a in agg.values()) / 1e9:.2f} |")
(REPO / "profiles").mkdir(exist_ok=True)
(REPO / "profiles" / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
gemm_bytes = sum(a[2] for k, a in agg.items() if k.startswith("gemm_bf16_nt"))
if "train" not in tag:   # bench.py's roofline.traffic is the FORWARD step's GEMM traffic
  (REPO / "profiles" / "gemm_traffic.json").write_text(json.dumps(
    {"dram_bytes_per_step": gemm_bytes, "source": f"profiles/{tag}_launches.md (ncu dram__bytes_read.sum + dram__bytes_write.sum "
     "summed over every gemm_bf16_nt_* launch of one forward step)"}, indent=1) + "\n")
print("\n".join(out))
