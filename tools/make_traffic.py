#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum per
launch of the stencil kernel in one bench step.  bench.py reports `roofline.traffic` from this file
only while the kernel source's sha256 still matches (so the figure cannot go stale silently).
`kernel` and `source.cu` may be comma-separated lists (the channelwise class runs two kernels).
usage: python tools/make_traffic.py gpurun_out/dw_traffic.csv <kernel> <source.cu> <workload> <clips> <launches_per_step>"""
import csv, hashlib, json, os, sys

path, kernel, source, workload, clips, nl = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
per = {}
for r in csv.DictReader(lines):
    if not any(k in r.get("Kernel Name", "") for k in kernel.split(",")):
        continue
    d = per.setdefault(int(r["ID"]), {})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    d[r["Metric Name"]] = v
ids = sorted(per)[-nl:]
rd = sum(per[i]["dram__bytes_read.sum"] for i in ids)
wr = sum(per[i]["dram__bytes_write.sum"] for i in ids)
h = hashlib.sha256()
for src in source.split(","):
    with open(os.path.join(root, src), "rb") as f:
        h.update(f.read())
sha = h.hexdigest()
entry = {"kernel": kernel, "workload": workload, "clips": clips, "launches": len(ids),
         "dram_bytes_read": rd, "dram_bytes_write": wr, "bytes_per_launch": (rd + wr) / len(ids),
         "source": source, "source_sha256": sha,
         "capture": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:%s on `bench.py --steps 1 "
                    "--warmup 3 --no-e2e --no-cpu-baseline --no-configs` (last %d launches = one step)" % (kernel.replace(",", "|"), nl)}
out = os.path.join(root, "profiles", "r02_traffic.json")
entries = []
if os.path.exists(out):
    entries = [e for e in json.load(open(out)) if not (e["kernel"] == kernel and e["workload"] == workload and e["clips"] == clips)]
entries.append(entry)
json.dump(entries, open(out, "w"), indent=1)
print(json.dumps(entry, indent=1))
