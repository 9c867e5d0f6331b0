"""Regenerates the golden fixtures in this directory from the read-only reference checkout.

Run in the build container only (`/root/reference` does not exist on the GPU box):
    python tests/golden/make_golden.py

Fixtures (all derived from data files the reference ships, no source code):
  summaries.json            parsed Keras `summary()` dumps  models/X3D-*/X3D_*.txt  (layer output
                            shapes, per-layer / total / trainable / non-trainable parameter counts)
  checkpoint_index.json     every key of models/X3D-{XS,S,M}/model.index with dtype enum, shape,
                            offset, size and masked CRC-32C (names/shapes/offsets are identical in
                            the three bundles; CRCs are per bundle)
  X3D-M.model.index         byte copy of models/X3D-M/model.index (61 kB table file; the data shard
                            is absent upstream) -- exercises the table reader on a real TF file
"""
import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference"

from x3d_tf_b200 import tf_bundle as tb  # noqa: E402


def parse_summary(path):
    rows, totals = [], {}
    with open(path) as f:
        for line in f:
            m = re.match(r"^(\S+) \((\w+)\)\s+\[?\((.*?)\)\]?\s+(\d+)\s*$", line)
            if m:
                shape = [None if s.strip() == "None" else int(s) for s in m.group(3).split(",")]
                rows.append({"name": m.group(1), "type": m.group(2), "shape": shape,
                             "params": int(m.group(4))})
            m = re.match(r"^(Total|Trainable|Non-trainable) params: ([\d,]+)", line)
            if m:
                totals[m.group(1).lower().replace("-", "_")] = int(m.group(2).replace(",", ""))
    return {"layers": rows, **totals}


def main():
    summ = {}
    for v in ("XS", "S", "M", "L", "XL"):
        summ[f"X3D_{v}"] = parse_summary(f"{REF}/models/X3D-{v}/X3D_{v}.txt")
    with open(os.path.join(HERE, "summaries.json"), "w") as f:
        json.dump(summ, f, indent=1)

    idx = {"keys": None, "crc": {}}
    for v in ("XS", "S", "M"):
        rd = tb.BundleReader(f"{REF}/models/X3D-{v}/model")
        table = [[k, e.dtype, list(e.shape), e.offset, e.size] for k, e in rd.entries.items()]
        if idx["keys"] is None:
            idx["keys"] = table
        else:
            assert idx["keys"] == table, "XS/S/M bundles are expected to share their layout"
        idx["crc"][f"X3D_{v}"] = [e.crc32c for e in rd.entries.values()]
    with open(os.path.join(HERE, "checkpoint_index.json"), "w") as f:
        json.dump(idx, f, separators=(",", ":"))
    shutil.copyfile(f"{REF}/models/X3D-M/model.index", os.path.join(HERE, "X3D-M.model.index"))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
