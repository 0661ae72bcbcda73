"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv
python bench.py ...`) per (kernel, grid): launches, average time, share of the profiled GPU time.

    python scripts/launch_summary.py profiles/r2_launches.csv > profiles/r2_launches_summary.txt
"""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, ig, iv, iu = (hdr.index(n) for n in ("Kernel Name", "Grid Size", "Metric Value", "Metric Unit"))
    acc = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        t = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        a = acc[(r[ik], r[ig])]
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in acc.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':<90} {'grid':>16} {'n':>5} {'avg_us':>10} {'share%':>7}")
    for (k, g), (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:90]:<90} {g:>16} {n:>5} {t / n:>10.2f} {100 * t / total:>7.1f}")
    # the single-solve step of the headline workload: march (1030 x 1 grid) + the two k_fft24 passes behind it
    step = {k: v for k, v in acc.items() if ("k_march" in k[0] and k[1].endswith(", 1, 1)")) or
            ("k_fft24<" in k[0] and k[1].endswith(", 2, 1)"))}
    if step:
        main_march = max((k for k in step if "k_march" in k[0]), key=lambda k: step[k][0])
        parts = {k: step[k][1] / step[k][0] for k in step if "k_fft24<" in k[0] or k == main_march}
        tot = sum(parts.values())
        print("# one single-solve step (config 2, the arithmetic mode the bench ran): kernel shares")
        for k, t in sorted(parts.items(), key=lambda kv: -kv[1]):
            print(f"#   {k[0][:80]:<80} {t:>8.2f} us  {100 * t / tot:>5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
