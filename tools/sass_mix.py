"""Dynamic opcode mix per kernel from `ncu -i rep --page source --csv --print-source sass` (Instructions Executed per SASS line).
Usage: python tools/sass_mix.py file.csv [top]"""
import csv, re, sys, collections
csv.field_size_limit(1 << 30)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kern = None; hdr = None; ops = None; done = set()
def flush():
    if kern is None or not ops: return
    tot = sum(ops.values())
    print("##", kern[:90], "executed warp instructions", tot)
    for op, n in ops.most_common(top): print(f"  {op:10s} {n:13d} {100 * n / tot:5.1f}%")
for r in csv.reader(open(sys.argv[1], errors="ignore")):
    if not r: continue
    if r[0] == "Kernel Name":
        flush(); kern = r[1]; ops = collections.Counter(); hdr = None
        if kern in done: kern = None
        else: done.add(kern)
        continue
    if r[0] == "Address": hdr = r; iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); continue
    if hdr is None or kern is None or len(r) <= iE: continue
    try: n = int(r[iE])
    except ValueError: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    if m: ops[m.group(2).split(".")[0]] += n
flush()
