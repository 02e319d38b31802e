"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv ...`)
per kernel: python tools/launch_summary.py x.csv "<command that was profiled>" > profiles/x_summary.txt"""
import csv
import re
import sys


def main():
    path, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "?")
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, total, n = {}, 0.0, 0
    for r in rows[1:]:
        if len(r) <= iv or "gpu__time_duration" not in ",".join(r):
            continue
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v
        name = re.sub(r"\(.*", "", r[ik])
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
        total += ms
        n += 1
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none of `{what}`")
    print(f"# ({path.split('/')[-1]}).  Cold-cache, serialised per-launch times: the SHARES are what is comparable with the live")
    print("# CUDA-event breakdowns (bench.py `roofline.share_of_kernel_time`, tools/profile_training.py).")
    print(f"# launches {n}, total {total:.1f} ms")
    for name, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:9.2f} ms {c:6d} launches {100 * ms / total:5.1f} %  {name}")


main()
