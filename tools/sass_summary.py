"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG, legacy mma.sync -> HMMA) in lpi_b200/liblpi_b200.so.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt        (cuobjdump from the CUDA toolkit; no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lpi_b200", "liblpi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDG.E.ENL2.256", "STG.E.ENL2.256",
        "MUFU.TANH", "MUFU.EX2", "ATOMG", "REDG"]
per = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        per[cur]["_instr"] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    per[cur]["_instr"] += 1
    for k in KEYS:
        if op == k or op.startswith(k + ".") or (k in ("LDG.E.ENL2.256", "STG.E.ENL2.256") and op.startswith(k)):
            if k == "UTCHMMA" and ".2CTA" in op:
                per[cur]["UTCHMMA.2CTA"] += 1
            elif k == "UTCHMMA.2CTA":
                continue
            else:
                per[cur][k] += 1
tot = collections.Counter()
print(f"# {os.path.relpath(lib, ROOT)}: {len(per)} kernels; columns = count of each mnemonic in the kernel's SASS (sm_100a)")
print("# kernel | instructions | " + " | ".join(KEYS))
for name, c in per.items():
    for k in KEYS:
        tot[k] += c[k]
    if not any(c[k] for k in KEYS if k not in ("SYNCS", "ATOMG", "REDG", "MUFU.EX2", "MUFU.TANH")):
        continue
    short = re.sub(r"\(.*", "", re.sub(r"\((int|bool|unsigned int)\)", "", demangle(name)))
    print(f"{short} | {c['_instr']} | " + " | ".join(str(c[k]) for k in KEYS))
print("# TOTAL over all kernels | - | " + " | ".join(str(tot[k]) for k in KEYS))
