"""SASS instruction histogram of the built library (profiles/r2_sass_histogram.md): the mnemonics that prove what the
kernels are made of (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, bulk
TMA copies -> UBLKCP, mma.sync -> HMMA, ...), per kernel.

    python scripts/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ikflow_b200", "lib", "libikflow_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UBLKPF", "STAS", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDSM", "MEMBAR", "FENCE", "UCGABAR",
       "ELECT", "FFMA2", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "LD.E", "ST.E", "ATOM", "RED", "BAR", "SHFL", "NANOSLEEP", "LDL", "STL"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEY:
            if op == k or op.startswith(k + ".") or (k == "UCGABAR" and op.startswith(k)):
                per[cur][k] += 1
                break


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


print(f"# SASS instruction histogram of `{os.path.relpath(lib)}` (`cuobjdump -sass`, sm_100a), per kernel")
print("# UTCHMMA = tcgen05.mma.kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (bulk TMA, incl. .multicast::cluster),")
print("# UBLKPF = cp.async.bulk.prefetch.L2, STAS = st.async (distributed shared memory hand-over of the k-split kernels), SYNCS = mbarrier ops, UCGABAR = barrier.cluster, HMMA = mma.sync (fallback engine only), LDL/STL = local memory\n")
cols = [k for k in KEY if any(c[k] for c in per.values())]
print("| kernel | instructions | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for name, c in per.items():
    short = demangle(name)
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
