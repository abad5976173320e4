"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel the
number of launches, mean and total device time and the share of all launches (dev aid)."""
import collections, csv, sys

def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            agg.setdefault(r[kn].split("(")[0], []).append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':66s} {'n':>4s} {'mean ms':>10s} {'total ms':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:66]:66s} {len(v):4d} {sum(v)/len(v)/1e6:10.4f} {sum(v)/1e6:10.3f} {100*sum(v)/tot:6.2f}%")

if __name__ == "__main__":
    main(sys.argv[1])
