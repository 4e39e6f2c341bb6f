"""SASS evidence of what the built library contains: per kernel, counts of the mnemonics that matter here (TMA bulk
tensor loads, mbarrier operations, 2-wide fp32, shared-memory traffic, shuffles).  Runs on the CPU (cuobjdump).

    python scripts/sass_counts.py [natrix_b200/libnatrix_b200.so] > profiles/<tag>_sass_counts.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "natrix_b200/libnatrix_b200.so"
COLS = ["UTMALDG", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FFMA", "FADD", "FMUL", "FSEL", "LDS", "STS", "SHFL", "LDG", "STG", "MUFU", "F2I"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kernels, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"natrix::\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        kernels[name][m.group(1)] += 1
        kernels[name]["_all"] += 1
print(f"# cuobjdump -sass {so}: architectures {arch}; static instruction counts per kernel")
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in COLS))
tot = collections.Counter()
for k, c in sorted(kernels.items()):
    print(f"{k[:58]:58s} {c['_all']:6d} " + " ".join(f"{c[x]:7d}" for x in COLS))
    tot.update(c)
print(f"{'TOTAL (' + str(len(kernels)) + ' kernels)':58s} {tot['_all']:6d} " + " ".join(f"{tot[x]:7d}" for x in COLS))
