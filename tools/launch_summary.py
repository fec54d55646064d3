"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (dev tool).

    python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_bench_launches.csv profiles/rNN_bench_launches_summary.json
"""
import collections
import csv
import json
import re
import sys


def main():
    src, out_csv, out_json = sys.argv[1:4]
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else (v * 1000 if row["Metric Unit"] in ("ms", "msecond") else v)
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        k = re.sub(r"<unnamed>::|\(anonymous namespace\)::|^void ", "", k)[:120]
        rows.append((row["ID"], k, row["Block Size"], row["Grid Size"], v))
        e = agg.setdefault(k, {"launches": 0, "us": 0.0})
        e["launches"] += 1
        e["us"] += v
    with open(out_csv, "w") as f:
        for h in sys.argv[4:]:
            f.write("# " + h + "\n")
        f.write("id,kernel,block,grid,duration_us\n")
        for r in rows:
            f.write(f'{r[0]},"{r[1]}",{r[2].replace(",", " ")},{r[3].replace(",", " ")},{r[4]:.3f}\n')
    tot = sum(e["us"] for e in agg.values())
    ordered = dict(sorted(agg.items(), key=lambda kv: -kv[1]["us"]))
    json.dump({"total_us": tot, "launches": len(rows), "kernels": ordered}, open(out_json, "w"), indent=0)
    for k, e in list(ordered.items())[:30]:
        print(f"{e['us']:10.1f} us {e['launches']:5d}  {100 * e['us'] / tot:5.1f}%  {k}")
    print("total", tot)


if __name__ == "__main__":
    main()
