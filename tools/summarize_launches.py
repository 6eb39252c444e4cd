"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total time, share.
  python tools/summarize_launches.py profiles/<launches>.csv [--alg]
With --alg the algorithmic bytes per launch (DESIGN.md section 4, 32^4) are put beside the measured time.
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys

NL = 4 * 32 ** 4
NS = 32 ** 4
ALG = {  # kernel-name fragment -> algorithmic bytes per launch at 32^4
    "lq_md4_kernel<128, 3, 1": 416 * NL, "lq_md4_kernel<128, 3, 0": 272 * NL, "KGaussField": 976 * NS, "lq_gfield4_kernel": 976 * NS, "lq_gstep4_kernel": 308 * NL, "lq_metro4_kernel": 144 * NL + 144 * NL // 8,
    "KGaussProjectStep": 308 * NL, "KPlaquette": 144 * NL, "lq_plaq4_kernel": 144 * NL, "KEfieldEnergy": 64 * NL, "KReunitarize": 288 * NL,
    "KGaussDiv": 144 * NS, "KMomentaRefresh": 64 * NL, "lq_sweep4_kernel": 144 * NL + 144 * NL // 8,
    "KHeatBath": 144 * NL + 144 * NL // 8, "KOverrelax": 144 * NL + 144 * NL // 8, "KMetropolis": 144 * NL + 144 * NL // 8,
}


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, newline="")))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r and "Metric Value" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        unit = r[hdr.index("Metric Unit")]
        try:
            v = float(r[hdr.index("Metric Value")].replace(",", ""))
        except ValueError:
            continue
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        key = re.sub(r"\(.*$", "", name).replace("void ", "").strip()
        n, t = agg.get(key, (0, 0.0))
        agg[key] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {total / 1e3:.2f} ms of kernel time")
    print(f"{'kernel':64s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>9s} {'alg GB/s':>9s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        alg = next((b for frag, b in ALG.items() if frag in k), None)
        gbs = f"{alg * n / (t * 1e-6) / 1e9:9.0f}" if alg and "--alg" in sys.argv else ""
        print(f"{k[:64]:64s} {n:8d} {t:12.1f} {100 * t / total:6.1f}% {t / n:9.1f} {gbs:>9s}")


if __name__ == "__main__":
    main()
