"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"])
        v = v / 1000 if d["Metric Unit"] == "ns" else v * 1000 if d["Metric Unit"] == "ms" else v
        a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%5s %12s %6s %10s  kernel" % ("n", "total_us", "share", "avg_us"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%5d %12.1f %5.1f%% %10.1f  %s" % (a[0], a[1], 100 * a[1] / tot, a[1] / a[0], k))
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
