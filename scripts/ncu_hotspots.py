#!/usr/bin/env python
"""Join an ncu report's per-instruction stall samples with source lines (-lineinfo).
usage: ncu_hotspots.py <report.ncu-rep> <lib.so> <substring of the MANGLED kernel name, e.g. 9k_eval_orILb0ELb0E> [top_n]
Prints the source lines with the most stall samples / executed instructions."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rows[0]["Address"], 16)
samples = [(int(r["Address"], 16) - base, int(r["# Samples"] or 0), int(r["Instructions Executed"] or 0), r["Source"].strip(), r) for r in rows]
# line table from the cubin
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
off2line = {}
for f in os.listdir(d):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    if kern not in txt:
        continue
    cur, infn, off = None, False, 0
    for l in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            infn = kern in m.group(1) and (("Lb0E" in m.group(1)) or True)
            off = 0
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            off2line.setdefault(int(m.group(1), 16), cur)
    break
agg = defaultdict(lambda: [0, 0])
tot_s = tot_i = 0
for off, s, n, src, _ in samples:
    k = off2line.get(off, ("?", 0))
    agg[k][0] += s
    agg[k][1] += n
    tot_s += s
    tot_i += n
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
srcs = {}
for (f, ln), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    text = ""
    for root in ("rucene_b200/csrc/gpu", "/usr/local/cuda/include", "/usr/local/cuda/targets/x86_64-linux/include/crt"):
        pth = os.path.join(root, f)
        if os.path.exists(pth):
            if pth not in srcs:
                srcs[pth] = open(pth, errors="replace").read().splitlines()
            if 0 < ln <= len(srcs[pth]):
                text = srcs[pth][ln - 1].strip()[:110]
            break
    print("%5.1f%% samples %5.1f%% instr  %s:%d  %s" % (100.0 * s / max(1, tot_s), 100.0 * n / max(1, tot_i), f, ln, text))
