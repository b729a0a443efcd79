"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time per evaluation."""
import collections
import csv
import sys

def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ki], float(r[vi].replace(",", ""))) for r in data if len(r) > vi]
    idx = [i for i, (k, _) in enumerate(seq) if "k_build(" in k or "k_chain_u" in k]
    if len(idx) < 2:
        print("fewer than two evaluations captured"); return
    n_ev = len(idx) - 1
    agg = collections.OrderedDict()
    for k, v in seq[idx[0]:idx[-1]]:
        name = k.split("(")[0]
        c = agg.setdefault(name, [0, 0.0]); c[0] += 1; c[1] += v
    tot = sum(v for _, v in agg.values())
    print(f"{n_ev} evaluations, {tot / 1000 / n_ev:.1f} us per evaluation (serialised, cold cache)")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:50s} launches/eval={c / n_ev:5.1f}  us/eval={v / 1000 / n_ev:9.2f}  share={v / tot:6.1%}  us/launch={v/1000/c:7.2f}")

if __name__ == "__main__":
    main(sys.argv[1])
