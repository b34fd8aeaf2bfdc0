"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel
count / total time / share (written to profiles/)."""
import csv
import collections
import re
import sys


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v * scale))
    agg = collections.OrderedDict()
    for n, us in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + us)
    tot = sum(t for _, t in agg.values())
    lines = [f"# per-kernel device time from {path} (ncu, cold-cache + serialised: compare SHARES)",
             f"# total {tot/1e3:.3f} ms over {len(rows)} launches", "kernel,launches,total_us,avg_us,share"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n},{c},{t:.1f},{t/c:.2f},{t/tot:.4f}")
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
