#!/usr/bin/env python
"""Per-kernel histogram of the Blackwell-specific SASS opcodes in the built library (cuobjdump -sass; no GPU needed):
UTCHMMA / UTCQMMA (tcgen05.mma kind::f16|tf32 / kind::f8f6f4, `.2CTA` = cta_group::2), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
HMMA (mma.sync), LDGSTS (cp.async), LDSM (ldmatrix).
  python profiles/sass_opcodes.py [lib.so] > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "gnn-lm_b200", "libgnnlm_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
OPS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "LDGSTS", "LDSM"]
pat = re.compile(r"\b(" + "|".join(OPS) + r")((?:\.[A-Z0-9_]+)*)")
per, arch, cur = collections.OrderedDict(), set(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur and "/*" in line:
        m = pat.search(line)
        if m:
            op, suf = m.group(1), m.group(2)
            key = op + (".2CTA" if ".2CTA" in suf else "") + (".MULTICAST" if "MULTICAST" in suf else "")
            per[cur][key] += 1
print(f"# {os.path.basename(lib)}: arch {sorted(arch)}; {len(per)} kernels; opcode counts per kernel (kernels without any of them omitted)")
tot = collections.Counter()
for fn, c in per.items():
    if not c:
        continue
    tot.update(c)
    name = demangle(fn)
    name = name[:name.index(">(") + 1] if ">(" in name else re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("gnnlm::", "")
    print(f"{name[:110]:110s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
print("# total: " + " ".join(f"{k}={v}" for k, v in sorted(tot.items())))
