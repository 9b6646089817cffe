"""Split an ncu capture of a warp-specialised kernel into barrier-delimited SASS segments.
usage: python scripts/ncu_segments.py report.ncu-rep  (needs ncu on PATH; no GPU)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; data = rows[2:]
iS, iE, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = 0; cur = 0; curS = 0; n = 0; st = {}
print("%-8s %-44s %12s %8s %6s  top stalls" % ("sass#", "ends with", "warp instr", "samples", "static"))
for k, r in enumerate(data):
    try: e = int(r[iE]); s = int(r[iS])
    except (ValueError, IndexError): continue
    tot += e; cur += e; curS += s; n += 1
    for i in stall_cols:
        try: st[hdr[i]] = st.get(hdr[i], 0) + int(r[i])
        except ValueError: pass
    if "BAR" in r[iSrc] or "EXIT" in r[iSrc]:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print("%-8d %-44s %12d %8d %6d  %s" % (k, r[iSrc].strip()[:44], cur, curS, n, " ".join("%s=%d" % (a[6:], b) for a, b in top)))
        cur = 0; curS = 0; n = 0; st = {}
print("total warp instructions", tot, "static", len(data))
