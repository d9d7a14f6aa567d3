#!/usr/bin/env python
"""Turns an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum` log of one bench step into
profiles/r02_predict_tc_traffic.json (read by bench.py's roofline.traffic):
    python tools/ncu_traffic_to_json.py gpurun_out/ncu_traffic_headline.csv headline [kernel-substring]"""
import csv
import json
import os
import subprocess
import sys

src, workload = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "predict_tc_kernel"
rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
per = {}
for r in rows:
    if kern not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "byte").lower()
    v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)
    per.setdefault(r["ID"], 0.0)
    per[r["ID"]] += v
# The kernel name covers three modes (Cholesky update, triangular inverse, predict GEMM) and the capture covers every step the
# bench runs (3 warm-up + 1 timed + 2 of the e2e leg).  Keep the TIMED step (the 4th) and, inside it, the predict-GEMM launches:
# the last `npredict` launches of the step (the factor-side modes come first).
steps, npredict = 6, int(sys.argv[4]) if len(sys.argv) > 4 else 4
ids = sorted(per, key=int)
per_step = len(ids) // steps
timed = ids[3 * per_step:4 * per_step][-npredict:]
per = {k: per[k] for k in timed}
launches = len(per)
total = sum(per.values())
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles", "r02_predict_tc_traffic.json")
d = json.load(open(out)) if os.path.exists(out) else {"workloads": {}}
d["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=root).stdout.strip()
d["how"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:%s on "
            "`python bench.py --no-next --no-cpu --steps 1 --warmup 3`; the launches of the last (timed) step" % kern)
d["workloads"][workload] = {"kernel": kern, "launches_per_step": launches, "dram_bytes_per_step": total,
                            "dram_bytes_per_launch": [per[k] for k in sorted(per, key=int)]}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d["workloads"][workload]))
