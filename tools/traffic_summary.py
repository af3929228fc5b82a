"""DRAM traffic per kernel family from an ncu launch list taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
of `tools/profile_step.py E workload 2`; the LAST step is summarised, cut at the marker launches
profile_step.py emits (tools/_ncu_csv.py) and checked to contain exactly one sgd_clip_update.
usage: python tools/traffic_summary.py launches.csv [out.json]"""
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _ncu_csv import last_step, load_launches  # noqa: E402


def main():
    path = sys.argv[1]
    launches = last_step(load_launches(path))
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in launches:
        k = re.sub(r"^void ", "", re.sub(r"\(.*", "", d["name"]))
        k = "itn::gemm_tf32_kernel" if "gemm_tf32_kernel" in k else k
        k = re.sub(r"attn_(\w+)_kernel<.*", r"attn_\1_kernel", k)[:60]
        v = fam[k]
        v[0] += 1
        v[1] += d.get("gpu__time_duration.sum", 0.0)
        v[2] += d.get("dram__bytes_read.sum", 0.0)
        v[3] += d.get("dram__bytes_write.sum", 0.0)
    tot_t = sum(v[1] for v in fam.values())
    tot_b = sum(v[2] + v[3] for v in fam.values())
    print(f"{len(launches)} launches, {tot_t/1e6:.2f} ms serialised, {tot_b/1e9:.2f} GB DRAM traffic per step")
    out = {"launches": len(launches), "step_ms_serialised": tot_t / 1e6, "dram_gb_per_step": tot_b / 1e9, "kernels": {}}
    for k, v in sorted(fam.items(), key=lambda kv: -(kv[1][2] + kv[1][3]))[:20]:
        gb = (v[2] + v[3]) / 1e9
        print(f"{gb:9.3f} GB (r {v[2]/1e9:8.3f} w {v[3]/1e9:8.3f}) {v[1]/1e6:8.3f} ms {gb/(v[1]/1e9) if v[1] else 0:8.0f} GB/s n={v[0]:4d}  {k}")
        out["kernels"][k] = {"launches": v[0], "ms": v[1] / 1e6, "dram_read_bytes": v[2], "dram_write_bytes": v[3]}
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
