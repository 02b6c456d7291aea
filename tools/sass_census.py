"""Counts the Blackwell-specific SASS mnemonics per kernel of libneuradar_b200.so (cuobjdump -sass): tcgen05.mma (UTC*MMA),
tensor-memory loads / stores (LDTM / STTM), TMA (UTMALDG tensor tiles, UBLKCP bulk copies), mbarrier waits (SYNCS) and the
vector reductions of the scatter kernels (RED / REDG), and the in-switch reduction loads of the peer all-reduce (LDGMC).  Usage: python tools/sass_census.py > profiles/r2_sass_census.md"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "neuradar_b200/lib/libneuradar_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([
    ("tcgen05.mma", r"\bUTC[A-Z]*MMA"), ("tcgen05.ld", r"\bLDTM"), ("tcgen05.st", r"\bSTTM"), ("tcgen05.cp/commit", r"\bUTCBAR|\bUTCCP"),
    ("TMA tensor (UTMALDG)", r"\bUTMALDG"), ("TMA bulk (UBLKCP)", r"\bUBLKCP"), ("mbarrier (SYNCS)", r"\bSYNCS"),
    ("RED", r"\bREDG?\b|\bRED\."), ("LDG", r"\bLDG\b|\bLDG\."), ("SHFL", r"\bSHFL"), ("elect", r"\bELECT"),
    ("NVLS multimem.ld_reduce (LDGMC)", r"\bLDGMC"), ("system-scope st / ld (STRONG.SYS)", r"STRONG\.SYS"),
])
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for name, pat in pats.items():
        if re.search(pat, line):
            kernels[cur][name] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print("# SASS census of libneuradar_b200.so (sm_100a), `python tools/sass_census.py`\n")
print("| kernel | " + " | ".join(pats) + " |")
print("|---|" + "---:|" * len(pats))
for (k, c), d in zip(kernels.items(), demangled):
    if not any(c[n] for n in list(pats)[:7]) and c["RED"] == 0 and c["NVLS multimem.ld_reduce (LDGMC)"] == 0:
        continue
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("nrb::", "")
    print(f"| `{short}` | " + " | ".join(str(c[n]) for n in pats) + " |")
