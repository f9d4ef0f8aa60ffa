"""Aggregates `ncu --page source --print-source cuda,sass --csv` output per CUDA source line (samples, executed
instructions) and prints the top lines. Usage: python tools/ncu_lines.py report.csv [top]"""
import csv, collections, glob, os, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csv.field_size_limit(1 << 30)
hdr = None; cur = None
by = collections.defaultdict(lambda: [0, 0])
for r in csv.reader(open(path, errors="ignore")):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": print("##", r[1][:100]); continue
    if r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= iE: continue
    try: ln = int(r[0]); s = int(r[iS] or 0); e = int(r[iE] or 0)
    except ValueError: continue
    if r[2] in ("", "-"): by[(cur, ln)][0] += s; by[(cur, ln)][1] += e
T = sum(v[0] for v in by.values()) or 1; E = sum(v[1] for v in by.values()) or 1
src = {}
for f in set(k[0] for k in by):
    for p in glob.glob(os.path.join(root, "lineslam_b200", "csrc", "**", f), recursive=True): src[f] = open(p).read().split("\n")
print("total samples", T, "executed", E)
for (f, ln), (s, e) in sorted(by.items(), key=lambda x: -x[1][0])[:top]:
    text = src[f][ln - 1].strip()[:100] if f in src and ln - 1 < len(src[f]) else ""
    print(f"{f}:{ln:4d} samples {100*s/T:5.2f}% exec {100*e/E:5.2f}%  {text}")
