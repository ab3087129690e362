"""SASS evidence per kernel of the shipped library: counts of the tensor-core / TMEM / TMA / cluster instructions
(cuobjdump -sass; no GPU needed).   python profiles/scripts/sass_summary.py > profiles/r2_sass_summary.txt
  UTCHMMA  tcgen05.mma       LDTM  tcgen05.ld (TMEM -> registers)    UTMALDG  cp.async.bulk.tensor (TMA load)
  UTCBAR   tcgen05.commit    SYNCS mbarrier ops                      UCGABAR  cluster barrier
  MUFU     SFU ops (ex2 / lg2 / rcp)   FFMA  fp32 FMA   HMMA  legacy mma.sync path (expected: 0)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = os.path.join(ROOT, "tsdiff_b200", "libtsdiff_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "MAPA", "MUFU", "FFMA",
        "HMMA", "ATOM", "RED."]
cur, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        sizes[cur] = 0
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        sizes[cur] += 1
        for k in KEYS:
            if (re.search(r"(?<![A-Z])HMMA", line) if k == "HMMA" else k in line):
                counts[cur][k] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# %s  (sm_100a, %d kernels)" % (os.path.relpath(lib, ROOT), len(counts)))
print("# kernel | SASS instructions | " + " ".join(KEYS))
tot = collections.Counter()
for (name, c), dm in zip(counts.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::|\(int\)", "", dm).replace("void ", "").split("(")[0]
    print("%-70s %6d | %s" % (short[:70], sizes[name], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
    tot.update(c)
print("TOTAL | " + " ".join("%s=%d" % (k, tot[k]) for k in KEYS))
