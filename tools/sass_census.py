"""Counts the Blackwell-specific SASS mnemonics per kernel of the shipped library (cuobjdump -sass):
UTCHMMA (tcgen05.mma), UTMALDG / UTMAPF (TMA tensor load / L2 prefetch), UBLKCP (cp.async.bulk), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32), and the plain tensor-core mnemonics of
older generations (HMMA / IMMA: must be zero).   python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "highres-net_b200", "csrc", "libhrn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "UTMASTG", "FFMA2", "FADD2", "FMUL2", "HMMA", "IMMA", "MEMBAR", "LDG", "STG"]
print("# SASS census of highres-net_b200/csrc/libhrn_b200.so (stamp %s)" % open(lib + ".stamp").read()[:16])
print("# targets:", " ".join(sorted(set(re.findall(r"sm_\d+a?", elf)))))
print("%-58s %7s " % ("kernel", "instr") + " ".join("%7s" % w for w in want))
tot = collections.Counter()
for block in re.split(r"\n\s+Function : ", sass)[1:]:
    name = block.split("\n")[0].strip()
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    short = re.sub(r"\(.*", "", demangled.replace("(anonymous namespace)::", "").replace("void ", "")).replace("hrn::", "")
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", block)
    c = collections.Counter(ops)
    row = [sum(v for k, v in c.items() if k.startswith(w)) for w in want]
    for w, v in zip(want, row): tot[w] += v
    print("%-58s %7d " % (short[:58], len(ops)) + " ".join("%7d" % v for v in row))
print("%-58s %7s " % ("TOTAL", "") + " ".join("%7d" % tot[w] for w in want))
