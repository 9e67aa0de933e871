"""Opcode histogram of the built library per kernel (runs without a GPU):

    python profiles/tools/sass_histogram.py > profiles/sass_opcodes_r2.txt
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "historymatching_b200/libhm_b200.so"
KEYS = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "UBLKCP", "STAS", "SYNCS", "UCGABAR_ARV", "LDGSTS", "ARRIVES", "LDS", "STS", "LDG",
        "STG", "BAR", "SHFL", "FSEL", "ISETP", "HMMA", "UTCHMMA", "LDTM"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
kernels, cur = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        cur["n"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "BAR" and op == "BAR") or (k == "UCGABAR_ARV" and op.startswith("UCGABAR")):
                cur[k] += 1
                break


def short(name):
    d = demangle(name)
    d = re.sub(r"^void ", "", d)
    d = d.replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")
    d = re.sub(r"\(.*$", "", d)
    return d.replace("hmsim::<unnamed>::", "").replace("<unnamed>::", "").replace("hmsim::", "")


tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print(f"SASS opcode histogram of {LIB} (cuobjdump -sass, sm_100a), built from HEAD with build.py")
print(f"whole library: {tot['n']} instructions in {len(kernels)} kernels")
print("  " + "  ".join(f"{k}={tot[k]}" for k in KEYS))
print("(DMMA = mma.sync f64 tensor path - tcgen05 has no FP64 kind; UBLKCP = cp.async.bulk; STAS = st.async to distributed shared memory;\n"
      " SYNCS = mbarrier; UCGABAR = cluster barrier; LDGSTS = cp.async; ARRIVES = cp.async completion on an mbarrier; no HMMA / UTC*MMA /\n"
      " LDTM: the whole path is FP64)\n")
for name, c in sorted(kernels.items(), key=lambda kv: -kv[1]["n"]):
    print(f"{short(name):<64s} n={c['n']:6d}  " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
