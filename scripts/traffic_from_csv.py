"""profiles/traffic.json entries from one-pass ncu CSVs (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum`) of ONE Gram call: all launches of the call are summed (programs with a folded phase run as
two launches per chunk of super-tiles).   usage: python scripts/traffic_from_csv.py FILE.csv:CONFIG@N ..."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        f, key = arg.rsplit(":", 1)
        rows = list(csv.reader(open(f)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        h = rows[hi]
        ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
        tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
        kernels, ids = {}, set()
        for r in rows[hi + 1:]:
            if len(r) <= vi or r[mi] not in tot:
                continue
            v = float(r[vi].replace(",", ""))
            if r[mi] == "gpu__time_duration.sum":
                v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e-6)
            else:
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
            tot[r[mi]] += v
            ids.add(r[0])
            name = r[ki].split("(")[0].replace("void unnamed>::", "").replace("void cnngp::<unnamed>::", "")
            kernels[name] = kernels.get(name, 0) + (1 if r[mi] == "gpu__time_duration.sum" else 0)
        n = max(1, len(ids))
        rd, wr = tot["dram__bytes_read.sum"], tot["dram__bytes_write.sum"]
        out[key] = {"dram_bytes_per_launch": (rd + wr) / n, "dram_bytes_per_gram": rd + wr, "dram_read": rd, "dram_write": wr,
                    "launches_per_gram": n, "kernel": ", ".join(f"{k} x{v}" for k, v in kernels.items()),
                    "gram_ms_under_ncu": tot["gpu__time_duration.sum"], "source": os.path.relpath(f, ROOT)}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
