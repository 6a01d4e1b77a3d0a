"""SASS opcode summary of libpyrhe_b200.so (cuobjdump -sass): per kernel, the instructions that show which hardware
path it takes -- UTCIMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor loads), LDGSTS (cp.async),
DMMA (fp64 mma.sync), SYNCS (mbarrier), REDG (red.global).

    python tools/sass_summary.py [out.json]
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyrhe_b200", "csrc", "libpyrhe_b200.so")
KEYS = ["UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "LDGSTS", "DMMA", "SYNCS", "REDG", "ATOMG", "PRMT", "POPC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["instructions"] += 1
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    cur[k] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    doc = {"library": os.path.relpath(LIB, ROOT), "arch": "sm_100a", "total": dict(total),
           "kernels": {k: dict(v) for k, v in per.items()}}
    text = json.dumps(doc, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")
    for k, v in per.items():
        hot = {x: v[x] for x in KEYS if v.get(x)}
        print(f"{k[:60]:60s} {v['instructions']:6d} {hot}")
    print("total", {x: total[x] for x in KEYS})


if __name__ == "__main__":
    main()
