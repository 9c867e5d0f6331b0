#!/usr/bin/env python
"""Hot spots of one launch of an .ncu-rep: opcode mix with stall samples and the most-sampled
SASS instructions with their dominant stall reason.
usage: python tools/ncu_hot.py report.ncu-rep [launch_index] [top_n]"""
import collections, csv, io, subprocess, sys

path = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, start = r, i + 1
        break
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seen, data, tot, st = set(), [], 0, 0
ops, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) < len(hdr) or r[0] in seen:
        continue
    seen.add(r[0])
    try:
        n, s = int(r[iexec]), int(r[isamp])
    except ValueError:
        continue
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    per = [(int(r[i]) if r[i].isdigit() else 0, hdr[i]) for i in stall_cols]
    for v, h in per:
        stalls[h] += v
    data.append((s, n, r[isrc], max(per)))
    tot += n; st += s; ops[op] += n; samp[op] += s
print("warp instructions", tot, "samples", st)
for h, v in stalls.most_common(10):
    print(f"  {h:28s} {100 * v / max(sum(stalls.values()), 1):5.1f}%")
for op, n in ops.most_common(16):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%  samples {100 * samp[op] / max(st, 1):5.1f}%")
for s, n, src, top in sorted(data, reverse=True)[:topn]:
    print(f"{100 * s / max(st, 1):5.2f}% {n:9d} {src[:84]:84s} {top[1][6:]}:{top[0]}")
