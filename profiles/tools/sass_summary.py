"""Static evidence from the built library, no GPU needed: per kernel the registers / stack / static shared memory that
`cuobjdump -res-usage` reports and counts of the SASS mnemonics that show how the kernel moves data (cp.async = LDGSTS,
32-byte stores = STG.E.ENL2.256, warp reductions = REDUX, shared atomics = ATOMS, funnel shifts = SHF, local-memory
traffic = LDL/STL).  Usage:  python profiles/tools/sass_summary.py [lib.so] [out.md] [name filter regex]
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "pcodec_b200/libcpcodec.so"
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/sass_summary.md"
flt = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None

MNEMONICS = ["LDGSTS", "STG.E.ENL2.256", "STG.E.128", "STG.E.64", "LDG.E.128", "LDG.E.64", "LDS", "STS", "ATOMS", "REDUX", "SHFL", "SHF", "LOP3", "BAR", "LDL", "STL", "UTMALDG", "UBLKCP", "SYNCS"]


def demangle(names):
    res = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    short = []
    for d in res[: len(names)]:
        m = re.match(r"(?:void )?pcob200::([\w]+)(<[^>]*>)?", d)
        short.append((m.group(1) + (m.group(2) or "")) if m else d[:60])
    return short


res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.split("\n"):
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        usage[cur] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT") and not kv.startswith("TEXTURE") and not kv.startswith("SURFACE") and not kv.startswith("SAMPLER"))
        cur = None

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
n_instr = collections.Counter()
cur = None
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        n_instr[cur] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                counts[cur][mn] += 1

names = sorted(usage)
short = demangle(names)
rows = []
for n, s in zip(names, short):
    if flt and not flt.search(s):
        continue
    u = usage[n]
    rows.append((s, u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"), u.get("LOCAL", "?"), n_instr[n], counts[n]))
with open(out, "w") as f:
    f.write(f"Static resource usage and SASS mnemonic counts of `{lib}` (cuobjdump -res-usage / -sass, sm_100a).\n")
    f.write("STACK is per-thread local memory reserved for arrays ptxas could not keep in registers or for calls; LDL/STL count the\ninstructions that touch it.  SHARED is static shared memory only (dynamic shared memory is set at launch).\n\n")
    f.write("| kernel | REG | STACK | SHARED | SASS instr | " + " | ".join(MNEMONICS) + " |\n")
    f.write("|---|---|---|---|---|" + "---|" * len(MNEMONICS) + "\n")
    for s, reg, stack, sh, loc, ni, c in rows:
        f.write(f"| `{s}` | {reg} | {stack} | {sh} | {ni} | " + " | ".join(str(c.get(mn, 0)) for mn in MNEMONICS) + " |\n")
print("wrote", out, len(rows), "kernels")
