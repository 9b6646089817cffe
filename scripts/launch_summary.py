"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches.csv)."""
import csv, re, sys
from collections import defaultdict


def short(name):
    m = re.match(r"void (?:sde4::)?(k_\w+)", name)
    if not m:
        return re.sub(r"\(.*", "", name.replace("void ", ""))[:60]
    k = m.group(1)
    if k in ("k_cost", "k_rollout"):
        model = re.search(r"(\w+Model)<", name)
        lanes = re.search(r"EvLT<(?:\(int\))?(\d+)", name)
        tail = re.search(r">, (?:\(bool\))?(\w+), (?:\(bool\))?(\w+)(?:, (?:\(int\))?(\d+))?>\(", name)
        ev = "L%s" % lanes.group(1) if lanes else ("L1p" if "Ev1P" in name else "L1")
        t = " ".join(x for x in (tail.groups() if tail else ()) if x is not None)
        return "%s %s %s [%s]" % (k, model.group(1) if model else "?", ev, t)
    return k


def main(path):
    tot, cnt = defaultdict(float), defaultdict(int)
    with open(path) as f:
        rows = [r for r in f if r.startswith('"')]
    for r in csv.DictReader(rows):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        tot[k] += float(r["Metric Value"]) / 1e3
        cnt[k] += 1
    s = sum(tot.values())
    print("# per-kernel totals of %s (%d launches, %.1f ms)" % (path, sum(cnt.values()), s / 1e3))
    for k in sorted(tot, key=lambda k: -tot[k])[:30]:
        print("%-70s n=%5d total %10.1f us (%4.1f%%) avg %9.1f us" % (k, cnt[k], tot[k], 100 * tot[k] / s, tot[k] / cnt[k]))


if __name__ == "__main__":
    main(sys.argv[1])
