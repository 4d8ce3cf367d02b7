"""Per-kernel SASS evidence of the tcgen05 / TMA path in the in-tree libhgk.so.
Counts, for every kernel in the sm_100a cubin: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk), UTMALDG / UTMASTG (cp.async.bulk.tensor), UTMAPF / UBLKPF (bulk prefetch), SYNCS (mbarrier),
FFMA / HFMA2 (SIMT math) and the instruction total.
usage: python tools/sass_summary.py [path/to/libhgk.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pose_adv_aug_b200", "libhgk.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, check=True).stdout.decode()
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)).encode(),
                       stdout=subprocess.PIPE, check=True).stdout.decode().splitlines()
COLS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"),
        ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("U*PF", r"\bU(TMA|BLK)PF"),
        ("SYNCS", r"\bSYNCS"), ("FFMA", r"\bFFMA"), ("total", r"^\s+/\*[0-9a-f]{4,}\*/\s")]
rows = []
cur = None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.OrderedDict((c, 0) for c, _ in COLS)
        nm = next(it)
        nm = re.sub(r"\(.*", "", nm).replace("void ", "").replace("hgk::", "")
        rows.append((nm, cur))
        continue
    if cur is None:
        continue
    for c, rx in COLS:
        if re.search(rx, line):
            cur[c] += 1
arch = re.findall(r"arch = (sm_\w+)", sass)
print("SASS summary of %s (%s), %d kernels" % (os.path.relpath(lib, ROOT), ", ".join(sorted(set(arch))), len(rows)))
print("%-58s" % "kernel" + "".join("%9s" % c for c, _ in COLS))
tot = collections.OrderedDict((c, 0) for c, _ in COLS)
for nm, r in sorted(rows, key=lambda x: (-x[1]["UTC*MMA"], -x[1]["UBLKCP"] - x[1]["UTMALDG"], x[0])):
    print("%-58s" % nm[:58] + "".join("%9d" % r[c] for c, _ in COLS))
    for c in tot:
        tot[c] += r[c]
print("%-58s" % "TOTAL" + "".join("%9d" % tot[c] for c, _ in COLS))
