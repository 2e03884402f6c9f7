"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
usage: python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n]"""
import csv, collections, sys

def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        seq.append((r[ki], v))
    return seq

if __name__ == "__main__":
    seq = load(sys.argv[1])
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    seq = seq[skip:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in seq:
        k = k.replace("<unnamed>::", "").replace("void ", "")[:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in seq)
    print("%d launches, %.1f us total" % (len(seq), tot))
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print("%10.1f us %5.1f%% %4d  %s" % (v, 100 * v / tot, n, k))
