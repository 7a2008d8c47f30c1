"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of total)."""
import collections
import csv
import io
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[row["Metric Unit"]]
        k = row["Kernel Name"][:90]
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot:.3f} ms total (cold-cache, serialised)")
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        print(f"{ms:10.3f} ms {100 * ms / tot:5.1f}%  n={n:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
