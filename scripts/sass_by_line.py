"""Static SASS statistics of one kernel by source line (nvdisasm -g -c output).
usage: python scripts/sass_by_line.py disasm.sass '<regex on the .text section name>' [--top N]
Most of the hot-path code is straight-line and executed once per time step, so the static count per line
is a good proxy for the dynamic one (loops: the prologue, the rolled Ev1 layers)."""
import collections, re, sys
path, pat = sys.argv[1], re.compile(sys.argv[2])
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
INT = ("IMAD", "IADD", "LEA", "ISETP", "SHF", "LOP3", "MOV", "SEL", "PRMT", "P2R", "R2P", "PLOP", "UMOV", "UIADD", "ULEA", "USHF", "ULOP", "UIMAD", "CS2R", "S2R", "VOTE", "IABS", "I2F", "F2I", "UISETP", "USEL", "R2UR")
inside = False; cur = ("?", 0); by = collections.defaultdict(collections.Counter); ops = collections.Counter()
for line in open(path):
    if line.startswith(".text."):
        inside = bool(pat.search(line)); continue
    if not inside: continue
    if line.startswith(".section") or line.startswith(".text"): inside = False; continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        by[cur][op] += 1; ops[op] += 1
tot = sum(ops.values())
print("total static instructions", tot)
print("by opcode:", ", ".join("%s %d" % kv for kv in ops.most_common(28)))
nint = sum(c for o, c in ops.items() if o.startswith(INT))
print("integer/address/predicate/move share: %.1f%%" % (100.0 * nint / tot))
rows = sorted(by.items(), key=lambda kv: -sum(kv[1].values()))
for (f, l), c in rows[:top]:
    n = sum(c.values()); ni = sum(v for o, v in c.items() if o.startswith(INT))
    print("%5d (%4.1f%%) int %4d | %s:%d | %s" % (n, 100.0 * n / tot, ni, f, l, " ".join("%s%d" % kv for kv in c.most_common(6))))
