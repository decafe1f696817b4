"""Per-kernel census of the SASS mnemonics that prove which hardware paths the built library uses (run anywhere the
library is built; no GPU needed):  python scripts/sass_census.py > profiles/<name>.txt
  UTCHMMA / UTCQMMA ... = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG = TMA bulk tensor loads,
  SYNCS = mbarrier, DMMA = FP64 tensor cores, HMMA = legacy mma.sync (should be absent), ACQBULK / UBLKCP = bulk copies."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mfbiclust_b200", "lib")
WANT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "HMMA",
        "IMMA", "DFMA", "REDUX", "ATOM", "RED", "MEMBAR", "LDGSTS", "SHFL", "BAR"]
out = []
for obj in sorted(f for f in os.listdir(LIBDIR) if f.endswith(".o")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIBDIR, obj)], capture_output=True, text=True).stdout
    kernel = None
    counts = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernel = re.sub(r"\(.*", "", kernel)
            counts[kernel] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and kernel:
            op = m.group(1)
            counts[kernel]["_total"] += 1
            if op in WANT:
                counts[kernel][op] += 1
    out.append("== %s" % obj)
    for k, c in counts.items():
        marks = " ".join("%s=%d" % (op, c[op]) for op in WANT if c[op])
        out.append("%-70s %6d instr  %s" % (k[:70], c["_total"], marks))
print("\n".join(out))
