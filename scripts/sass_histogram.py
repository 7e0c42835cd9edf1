"""SASS evidence for profiles/: per-kernel instruction histogram of the built library (no GPU needed).

    python scripts/sass_histogram.py [path/to/libimpala_b200.so] > profiles/r2_sass_histogram.txt

Mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy),
SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32, UBLKPF = cp.async.bulk.prefetch.L2, ACQBULK / CCTL etc. as they appear."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "torched_impala_b200", "lib", "libimpala_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU",
       "LDG.E.128", "STG.E.128", "LDGSTS", "HMMA", "DADD", "DFMA", "SHFL", "BAR.SYNC", "ATOM", "RED", "MEMBAR", "LDS", "STS",
       "ELECT", "ACQBULK", "ERRBAR", "NANOSLEEP", "UBLKPF")
cur, hist, total, order = None, collections.defaultdict(collections.Counter), collections.Counter(), []
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEY:
            if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                hist[cur][k] += 1
print(f"# {os.path.relpath(lib, HERE)}: SASS instruction histogram per kernel (cuobjdump -sass)")
for name in order:
    h = hist[name]
    tags = " ".join(f"{k}={h[k]}" for k in KEY if h[k])
    print(f"{name}\n    total={total[name]}  {tags}")
