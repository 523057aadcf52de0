"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`)
into the per-kernel table committed under profiles/:
    python profiles/launch_list.py gpurun_out/launches.csv "command that was profiled" > profiles/launches_rNN.md"""
import csv
import sys
from collections import OrderedDict


def main(path, command):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    head = rows[0]
    iname, ival, iunit = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[head.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iunit], 1e-3)
        name = r[iname].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print("# Launch list, aggregated (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
    print(f"Command: `{command}`; per-launch times are cold-cache and serialised, compare shares.\n")
    print("| kernel | launches | total us | share | us/launch |\n|---|---|---|---|---|")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {c} | {t:.1f} | {100 * t / total:.1f}% | {t / c:.1f} |")
    print(f"\nTotal {total:.0f} us over {n} launches.")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
