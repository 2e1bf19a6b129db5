"""Per-kernel SASS mnemonic counts of the shipped library (the evidence table B200_PROFILING.md asks for):
python tools/sass_table.py > profiles/r02_sass_table.txt   (needs cuobjdump; no GPU)."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "humanoid-vision-system_b200", "libhvs_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "FMUL2", "MUFU", "LDGSTS", "total"]
pat = {"UTCHMMA": r"\bUTC[A-Z]*MMA", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG", "UBLKCP": r"\bUBLKCP", "LDTM": r"\bLDTM", "STTM": r"\bSTTM",
       "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "HMMA": r"\bHMMA", "FFMA2": r"\bFFMA2", "FMUL2": r"\bFMUL2", "MUFU": r"\bMUFU", "LDGSTS": r"\bLDGSTS"}
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        counts[kern]["total"] += 1
        for c, p in pat.items():
            if re.search(p, line):
                counts[kern][c] += 1
def short(name):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"hvs::\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d)
    return d[:58]
print(f"SASS mnemonic counts per kernel of {os.path.basename(lib)} (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA tensor load/store,")
print("UBLKCP = cp.async.bulk, LDTM/STTM = tcgen05.ld/st, HMMA = legacy mma.sync, FFMA2/FMUL2 = packed fp32x2")
print(f"{'kernel':58s} " + " ".join(f"{c:>7s}" for c in cols))
for k, c in counts.items():
    print(f"{short(k):58s} " + " ".join(f"{c[x]:7d}" for x in cols))
