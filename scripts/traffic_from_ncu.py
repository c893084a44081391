"""profiles/traffic.json from ncu summaries: DRAM bytes per launch of the Gram kernels at the sizes bench.py
runs them (`roofline.traffic`).   usage: python scripts/traffic_from_ncu.py SUMMARY.json:CONFIG@N ..."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(s):
    v, u = s.split()[:2]
    return float(v) * UNIT[u]


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        f, key = arg.rsplit(":", 1)
        d = json.load(open(f))
        rd, wr = num(d["dram__bytes_read.sum"]), num(d["dram__bytes_write.sum"])
        out[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                    "kernel": re.sub(r"\(.*", "", d["kernel"]).replace("void unnamed>::", ""),
                    "launch_ms_under_ncu": float(d["gpu__time_duration.sum"].split()[0]),
                    "source": os.path.relpath(f, ROOT)}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
