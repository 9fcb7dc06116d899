"""SASS evidence that the kernels are Blackwell-native: per kernel of libmvsnet_b200.so, the number of tcgen05 MMA
(UTC*MMA), TMEM load / store (LDTM / STTM), TMA tensor load (UTMALDG), bulk copy (UBLKCP) and mbarrier (SYNCS)
instructions.     python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "scene_3dreconstruction_mvsnet_b200", "libmvsnet_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                                ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("HFMA2", r"\bHFMA2"), ("FFMA2", r"\bFFMA2"), ("HMMA/MMA.SYNC", r"\bHMMA")])
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["instr"] += 1
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS instruction counts per kernel of libmvsnet_b200.so (cuobjdump -sass, sm_100a)\n")
print("| kernel | instr | " + " | ".join(pats) + " |")
print("|---|---|" + "---|" * len(pats))
tot = collections.Counter()
for (name, c), d in zip(counts.items(), demangled):
    short = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")).replace("void ", "").replace("mvs::", "")
    print("| `%s` | %d | " % (short[:70], c["instr"]) + " | ".join(str(c[k]) for k in pats) + " |")
    tot.update(c)
print("| **total** | %d | " % tot["instr"] + " | ".join(str(tot[k]) for k in pats) + " |")
