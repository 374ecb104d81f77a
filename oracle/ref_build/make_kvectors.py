#!/usr/bin/env python3
"""Writes kVectors3D.json / kNorm3D.csv in the format the reference's InitSystem() reads
(ReadKValuesFromJsonFile, src/Utils.cpp:1020-1049; BosonsBulk.cpp:124-137): integer wave vectors
grouped by shells of equal norm.  They only feed the S(k) observable (outside the hot path), but
the reference cannot initialise a system without them, and /root/reference is not available on the
GPU box.  Generated from first principles -- nothing is copied from the reference's config/."""
import json
import math
import os
import sys

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
shells = {}
m = 21
for x in range(m + 1):
    for y in range(m + 1):
        for z in range(m + 1):
            n2 = x * x + y * y + z * z
            if 0 < n2 <= m * m:
                shells.setdefault(n2, []).append([x, y, z])
keys = sorted(shells)[:400]
with open(os.path.join(out, "kVectors3D.json"), "w") as f:
    json.dump({"data": [sorted(shells[k]) for k in keys]}, f)
with open(os.path.join(out, "kVectors.json"), "w") as f:   # the name HeBulk / HeDrop read (HeBulk.cpp:139)
    json.dump({"data": [sorted(shells[k]) for k in keys]}, f)
with open(os.path.join(out, "kNorm3D.csv"), "w") as f:
    f.write("\n".join(repr(math.sqrt(k)) for k in keys) + "\n")

# one-dimensional shells for the 1-D systems (InhContactBosons reads kVectors1D.json / kNorm1D.csv, InhContactBosons.cpp:170-171)
with open(os.path.join(out, "kVectors1D.json"), "w") as f:
    json.dump({"data": [[[k]] for k in range(1, 401)]}, f)
with open(os.path.join(out, "kNorm1D.csv"), "w") as f:
    f.write("\n".join(repr(float(k)) for k in range(1, 401)) + "\n")

# two-dimensional shells (BosonsBulk / NUBosonsBulkPB with DIM = 2 read kVectors2D.json / kNorm2D.csv)
shells2 = {}
for x in range(m + 1):
    for y in range(m + 1):
        n2 = x * x + y * y
        if 0 < n2 <= m * m:
            shells2.setdefault(n2, []).append([x, y])
keys2 = sorted(shells2)[:120]
with open(os.path.join(out, "kVectors2D.json"), "w") as f:
    json.dump({"data": [sorted(shells2[k]) for k in keys2]}, f)
with open(os.path.join(out, "kNorm2D.csv"), "w") as f:
    f.write("\n".join(repr(math.sqrt(k)) for k in keys2) + "\n")
