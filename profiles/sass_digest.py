"""SASS digest of libplume_b200.so: per kernel, how many of the instructions that prove a Blackwell-native path
(profiling guide, "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UBLKCP / UTMALDG / UTMASTG = TMA (cp.async.bulk / .tensor load / .tensor store), HMMA = mma.sync, LDGSTS = cp.async, SYNCS = mbarrier,
FFMA2 / FMUL2 / FADD2 = packed fp32 pairs, MUFU.
    python profiles/sass_digest.py > profiles/r2_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "uav-wrf-les-ppo-lstm_b200", "build")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "LDGSTS", "SYNCS", "BAR",
        "FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "MUFU", "IMAD", "LDG", "STG", "LDS", "STS", "ATOM", "RED"]


def main():
    rows = []
    for obj in sorted(os.listdir(OBJ)):
        if not obj.endswith(".o"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        fn, counts, total = None, None, 0
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if fn:
                    rows.append((obj, fn, total, counts))
                fn, counts, total = m.group(1), collections.Counter(), 0
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
            if m and fn:
                total += 1
                op = m.group(1)
                for k in KEYS:
                    if op == k or (k in ("UTCHMMA", "UTCQMMA", "UTCBAR", "UBLKCP", "UTMALDG") and op.startswith(k)):
                        counts[k] += 1
        if fn:
            rows.append((obj, fn, total, counts))
    demangle = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("SASS digest of the kernels in libplume_b200.so (cuobjdump -sass, sm_100a); columns = static instruction counts")
    tot = collections.Counter()
    for (obj, fn, total, counts), name in zip(rows, demangle):
        name = re.sub(r"\(.*", "", name).replace("plume::", "")
        sel = "  ".join(f"{k}={counts[k]}" for k in KEYS if counts[k])
        print(f"{obj[:-2]:22s} {name[:60]:60s} instr={total:6d}  {sel}")
        tot.update(counts)
    print("TOTAL  " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    sys.exit(main())
