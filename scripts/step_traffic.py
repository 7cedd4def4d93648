"""profiles/<name>_ncu_step_launches.csv (ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
of scripts/ncu_step.py) -> profiles/<name>_step_traffic.json: per kernel, launches / time / DRAM bytes of one step.
    python scripts/step_traffic.py profiles/r1_g_ncu_step_launches.csv profiles/r1_g_step_traffic.json"""
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
per_id = {}
for row in csv.reader(open(src)):
    if len(row) < 15 or not row[0].isdigit():
        continue
    name = re.sub(r"^void ", "", row[4])
    name = re.sub(r"^avdf::", "", name)
    name = re.sub(r"[<(].*$", "", name)
    d = per_id.setdefault(row[0], {"name": name})
    val = float(row[14].replace(",", ""))
    unit = row[13]
    metric = row[12]
    if metric == "gpu__time_duration.sum":
        d["us"] = val / 1e3 if unit in ("nsecond", "ns") else (val if unit in ("usecond", "us") else val * 1e3)
    elif metric.startswith("dram__bytes"):
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d["rd" if "read" in metric else "wr"] = val * mult
kernels = {}
for d in per_id.values():
    k = kernels.setdefault(d["name"], {"launches_per_step": 0, "us_per_step": 0.0, "dram_read_bytes_per_step": 0.0, "dram_write_bytes_per_step": 0.0})
    k["launches_per_step"] += 1
    k["us_per_step"] += d.get("us", 0.0)
    k["dram_read_bytes_per_step"] += d.get("rd", 0.0)
    k["dram_write_bytes_per_step"] += d.get("wr", 0.0)
out = {"source": "ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "python scripts/ncu_step.py (one eager step, audio workload, batch 32, mixed precision, one lane): " + src,
       "kernels": dict(sorted(kernels.items(), key=lambda kv: -kv[1]["us_per_step"]))}
json.dump(out, open(dst, "w"), indent=1)
tot = sum(k["us_per_step"] for k in kernels.values())
for n, k in out["kernels"].items():
    print("%-34s n=%3d  %8.1f us (%4.1f%%)  DRAM rd %7.1f MB  wr %7.1f MB" % (n, k["launches_per_step"], k["us_per_step"], 100 * k["us_per_step"] / tot,
                                                                           k["dram_read_bytes_per_step"] / 1e6, k["dram_write_bytes_per_step"] / 1e6))
print("total %.1f us" % tot)
