"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown tables of time share per kernel.

    python scripts/summarize_launches.py gpurun_out/launches.csv [--last-steps 8 --per-step 47] > profiles/rNN_launches_summary.md
"""
import argparse
import csv
import re
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("__nv_bfloat16", "bf16")
    return name.replace("void ", "").strip()


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1.0)
        rows.append((short(r[ik]), ns))
    return rows


def table(rows, title):
    agg = OrderedDict()
    for k, ns in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    out = [f"## {title}", "", "| share | launches | avg us | kernel |", "|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {100 * ns / tot:.1f}% | {n} | {ns / n / 1e3:.1f} | `{k}` |")
    out += ["", f"Total {tot / 1e6:.1f} ms over {len(rows)} launches.", ""]
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--title", default="Whole call")
    ap.add_argument("--last-steps", type=int, default=0)
    ap.add_argument("--per-step", type=int, default=47)
    ap.add_argument("--tail-skip", type=int, default=0, help="launches after the last decode step (e.g. none)")
    a = ap.parse_args()
    rows = load(a.csv)
    print(table(rows, a.title))
    if a.last_steps:
        n = a.last_steps * a.per_step
        end = len(rows) - a.tail_skip
        print(table(rows[end - n:end], f"Last {a.last_steps} decode steps, {a.per_step} launches per step"))


if __name__ == "__main__":
    main()
