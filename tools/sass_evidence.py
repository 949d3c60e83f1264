#!/usr/bin/env python
"""Count the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md, "What proves a Blackwell-native kernel")
in every kernel of libvmm_sm100.so.  Runs without a GPU: cuobjdump -sass.  usage: tools/sass_evidence.py > profiles/r1_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "videometamaterials_b200", "libvmm_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDSM", "LDGSTS", "MUFU.TANH", "REDG", "ELECT"]
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void ", "").replace("vmm::", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in KEYS:
        if re.search(r"\b" + re.escape(k), line):
            per[cur][k] += 1
print("SASS mnemonics per kernel of `libvmm_sm100.so` (`cuobjdump -sass`, sm_100a; counts of static instructions).  "
      "`UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` = TMA tensor loads / stores, `UTCBAR` = tcgen05.commit, "
      "`SYNCS` = mbarrier, `HMMA` + `LDSM` = mma.sync + ldmatrix (attention cores), `LDGSTS` = cp.async, `ELECT` = elect.sync "
      "(one issuing lane), `REDG` = global reductions (no-return atomics; `F32x4` = the vector form), `MUFU.TANH` = one-instruction sigmoid.\n")
print("| kernel | " + " | ".join(KEYS) + " |")
print("|---|" + "---:|" * len(KEYS))
for name, c in per.items():
    if sum(c.values()) == 0:
        continue
    print(f"| `{name[:70]}` | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
