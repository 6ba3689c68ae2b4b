#!/usr/bin/env python
"""SASS instruction counts per tcgen05 kernel of the in-tree library (TMA / tensor-core / TMEM mnemonics, generic vs
shared-space memory instructions): python scripts/sass_counts.py > profiles/sass_r2.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "sp-gan_b200", "libspgan_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("Function : ")
keys = ["UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "STTM", r"\bLDS", r"\bSTS", r" LD\.E", r" ST\.E", "R2UR", "SYNCS"]
print("cuobjdump -sass %s (sm_100a); counts of SASS mnemonics per kernel" % os.path.relpath(lib, ROOT))
print("%-58s" % "kernel" + "".join("%9s" % k.replace("\\b", "").replace("\\", "").strip() for k in keys))
tot = dict.fromkeys(keys, 0)
for f in txt[1:]:
    name = f.split("\n")[0]
    if "UTCHMMA" not in f and "UTMALDG" not in f:
        continue
    short = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(anonymous namespace\)::", "", short).split("(")[0][:56]
    counts = [len(re.findall(k, f)) for k in keys]
    for k, c in zip(keys, counts):
        tot[k] += c
    print("%-58s" % short + "".join("%9d" % c for c in counts))
print("%-58s" % "total" + "".join("%9d" % tot[k] for k in keys))
