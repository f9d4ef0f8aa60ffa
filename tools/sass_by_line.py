"""Per CUDA source line: executed warp instructions split into fp64 / lds+sts / ldg+stg+local / int+addr / select+setp / branch / other.
Input: `ncu -i rep --page source --csv --print-source cuda,sass`. Usage: python tools/sass_by_line.py file.csv "kernel substring" [top]"""
import csv, re, sys, collections, os, glob
csv.field_size_limit(1 << 30)
want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
CLS = {"fp64": ("DMUL", "DADD", "DFMA", "DSETP", "MUFU"), "smem": ("LDS", "STS"), "gmem": ("LDG", "STG", "LDL", "STL", "LDC", "LDCU"),
       "int": ("IMAD", "IADD3", "VIADD", "LEA", "SHF", "LOP3", "UIMAD", "UIADD3", "ULEA", "UMOV", "MOV", "CS2R", "S2R", "S2UR", "PRMT", "IABS", "I2F", "F2I", "UFLO", "ULOP3"),
       "sel": ("FSEL", "SEL", "ISETP", "FSETP", "PLOP3", "UISETP", "VIMNMX", "FMNMX", "USEL"), "bra": ("BRA", "BSSY", "BSYNC", "WARPSYNC", "CALL", "RET", "BREAK", "EXIT", "NOP", "BAR", "SHFL", "VOTE")}
def cls(op):
    for k, v in CLS.items():
        if op in v: return k
    return "other"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fn = None; cur = None; hdr = None; line = None
by = collections.defaultdict(collections.Counter)
for r in csv.reader(open(sys.argv[1], errors="ignore")):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; iA = 2; iS = 3; iE = hdr.index("Instructions Executed"); continue
    if hdr is None or fn is None or want not in fn or len(r) <= iE: continue
    if r[0] != "": line = (cur, int(r[0])); continue
    try: n = int(r[iE])
    except ValueError: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    if m and line: by[line][cls(m.group(2).split(".")[0])] += n
tot = collections.Counter()
for c in by.values(): tot.update(c)
T = sum(tot.values())
print("total", T, {k: f"{100 * v / T:.1f}%" for k, v in tot.most_common()})
src = {}
for f in set(k[0] for k in by):
    for p in glob.glob(os.path.join(root, "lineslam_b200", "csrc", "**", f), recursive=True): src[f] = open(p).read().split("\n")
for (f, ln), c in sorted(by.items(), key=lambda x: -sum(x[1].values()))[:top]:
    n = sum(c.values())
    text = src[f][ln - 1].strip()[:70] if f in src and ln - 1 < len(src[f]) else ""
    print(f"{f}:{ln:4d} {100 * n / T:5.2f}%  " + " ".join(f"{k}={100 * c[k] / T:4.1f}" for k in ("fp64", "smem", "gmem", "int", "sel", "bra", "other") if c[k]) + "  | " + text)
