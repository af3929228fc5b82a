"""Shared by launch_summary.py / traffic_summary.py: read an ncu --csv launch list and cut it into the
steps `tools/profile_step.py` ran.  profile_step.py brackets every step with one launch of
`round_tf32_kernel` on a single element (a kernel the tf32x3 step itself never launches), so the step
boundaries are exact - the first step carries several hundred one-time launches (weight preparation, BN
folding) and "the second half of the list" is NOT the second step."""
import collections
import csv

MARKER = "round_tf32_kernel"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3,
        "nsecond": 1.0, "msecond": 1e6, "second": 1e9}


def load_launches(path):
    """-> list of {name, grid, <metric>: value (bytes / ns)} in launch order."""
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    return list(per.values())


def split_steps(launches):
    """Segments between consecutive marker launches (markers excluded)."""
    idx = [i for i, d in enumerate(launches) if MARKER in d["name"]]
    if len(idx) < 2:
        raise SystemExit(f"need at least two `{MARKER}` step markers in the launch list, found {len(idx)}: "
                         "take it with the current tools/profile_step.py")
    return [launches[a + 1:b] for a, b in zip(idx[:-1], idx[1:])]


def last_step(launches):
    steps = split_steps(launches)
    step = steps[-1]
    n_sgd = sum("sgd_clip_update" in d["name"] for d in step)
    assert n_sgd == 1, f"a predict() step runs sgd_clip_update exactly once, this segment has {n_sgd}"
    return step
