"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py launches.csv [--gemm] [--step LAST|ALL]"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _ncu_csv import last_step, load_launches  # noqa: E402


def main():
    path = sys.argv[1]
    launches = load_launches(path)
    if "--all" not in sys.argv:
        launches = last_step(launches)      # the last profiled step, cut at profile_step.py's marker launches
    rows = [{"Kernel Name": d["name"], "Grid Size": d["grid"], "t": d.get("gpu__time_duration.sum", 0.0)}
            for d in launches]
    dur = lambda r: r["t"]
    tot = sum(dur(r) for r in rows)
    print(f"{len(rows)} launches, {tot/1e6:.3f} ms (serialised, cold cache: compare shares)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"])
        k = re.sub(r"^void ", "", k)[:70]
        agg[k][0] += 1
        agg[k][1] += dur(r)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{v[1]/1e6:10.3f} ms {v[0]:5d} {100*v[1]/tot:6.1f}%  {k}")
    if "--gemm" in sys.argv:
        g = collections.defaultdict(lambda: [0, 0.0])
        for r in rows:
            m = re.search(r"gemm_tf32_kernel<(\d+), (\d), (\d), (\d)>", r["Kernel Name"])
            if m:
                key = (int(m.group(1)), m.group(2) + m.group(3), r["Grid Size"])
                g[key][0] += 1
                g[key][1] += dur(r)
        print("\nGEMM launches by (BN, majors A/B, grid):")
        for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:45]:
            print(f"{v[1]/1e6:8.3f} ms n={v[0]:3d} avg {v[1]/v[0]/1e3:8.1f} us  BN={k[0]:3d} majors={k[1]} grid={k[2]}")


if __name__ == "__main__":
    main()
