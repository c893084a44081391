"""Per-kernel SASS opcode histogram of libcnngp.so (cuobjdump -sass), the evidence that the hot
kernels are sm_100a-native: FFMA2 / FADD2 / FMUL2 (packed f32x2), MUFU, UBLKCP (cp.async.bulk),
SYNCS (mbarrier), LDTM / STTM (tcgen05.ld / st), USETMAXREG, DMMA (fp64 tensor pipe).
usage: python scripts/sass_histogram.py [lib.so] [out.json]   (runs without a GPU)"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "FMNMX", "LDS", "STS", "UBLKCP", "SYNCS", "LDTM", "STTM",
        "UTCBAR", "USETMAXREG", "DMMA", "DFMA", "DADD", "DMUL", "LDGSTS", "LDG", "STG", "SHFL", "BAR", "LDL", "STL", "BRX", "ATOMG", "REDG")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cnn-gp_b200", "libcnngp.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|cnngp::|\(int\)", "", name).split("(")[0].replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    out = {}
    for name, c in sorted(kernels.items(), key=lambda kv: -kv[1]["_total"]):
        out[name] = {"instructions": c["_total"], **{k: c[k] for k in KEEP if c[k]}}
    s = json.dumps({"library": os.path.relpath(lib, ROOT), "arch": "sm_100a", "kernels": out}, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s + "\n")
    for name, d in out.items():
        print(f"{d['instructions']:>7}  {name[:70]:70s} " + " ".join(f"{k}={v}" for k, v in d.items() if k != "instructions" and k in
              ("FFMA2", "FADD2", "MUFU", "UBLKCP", "SYNCS", "LDTM", "STTM", "USETMAXREG", "DMMA", "LDL", "STL")))


if __name__ == "__main__":
    main()
