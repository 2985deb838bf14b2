#!/usr/bin/env python
"""Per-kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file <list.csv> ...`).

    python tools/launch_shares.py gpurun_out/r2_launches.csv profiles/r2_launch_shares.csv
"""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
    tot = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*$", "", r[ik]).strip()
        us = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + us)
    total = sum(t for _, t in tot.values())
    with open(dst, "w") as f:
        f.write("# launch list of `python bench.py --steps 2 --warmup 3 --no-secondary` (ncu --metrics gpu__time_duration.sum "
                "--clock-control none; cold-cache, serialised: compare SHARES)\n")
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct"])
        for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, round(t, 1), round(100.0 * t / total, 2)])
        f.write(f"# total {total:.1f} us over {sum(n for n, _ in tot.values())} launches\n")


if __name__ == "__main__":
    main()
