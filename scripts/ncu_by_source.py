"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by source file / function / line.
usage: python scripts/ncu_by_source.py export.csv [csrc_dir]"""
import csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sde4mbrl_b200", "csrc")
hdr = None; fname = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10 or r[0] == "": continue
    try: s = int(r[6]); ie = int(r[7])
    except ValueError: continue
    out.append((s, ie, fname, int(r[0]), r[1]))
tot = sum(o[0] for o in out); toti = sum(o[1] for o in out)
print("samples", tot, "warp instructions", toti)
byf = {}
for o in out:
    a = byf.setdefault(o[2], [0, 0]); a[0] += o[0]; a[1] += o[1]
for f, a in sorted(byf.items(), key=lambda kv: -kv[1][1]): print("%-28s %5.1f%% smp %5.1f%% ins" % (f, 100 * a[0] / tot, 100 * a[1] / toti))
def funcs(path):
    res = []; cur = "?"
    for l in open(path):
        m = re.match(r"\s*(?:template.*>\s*)?(?:SDE4_DEV|__global__|static|__device__)[^;(]*?\b(\w+)\s*\(", l)
        if m and not l.strip().startswith("//"): cur = m.group(1)
        res.append(cur)
    return res
agg = {}
for f in byf:
    p = os.path.join(src, f or "")
    if not f or not os.path.exists(p): continue
    fm = funcs(p)
    for o in out:
        if o[2] == f:
            k = (f, fm[o[3] - 1] if o[3] - 1 < len(fm) else "?")
            a = agg.setdefault(k, [0, 0]); a[0] += o[0]; a[1] += o[1]
print()
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:50]: print("%-18s %-26s %5.1f%% smp %5.1f%% ins" % (k[0], k[1], 100 * a[0] / tot, 100 * a[1] / toti))
if "--lines" in sys.argv:
    out.sort(key=lambda o: -o[0])
    for o in out[:60]: print("%5.2f%% smp %5.2f%% ins | %s:%d | %s" % (100 * o[0] / tot, 100 * o[1] / toti, o[2], o[3], o[4].strip()[:100]))
