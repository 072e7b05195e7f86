"""Per-kernel SASS mnemonic counts of the objects that make up libnerfb200.so (cuobjdump -sass, sm_100a): the evidence
that the tensor-core / TMEM / bulk-copy paths are what the kernels execute.  python scripts/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "nerf_rep_for_test_b200", "build")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "SYNCS", "FFMA2", "FADD2", "FFMA", "DFMA", "DMUL", "MUFU", "REDG",
        "ATOMG", "F2FP"]
print("SASS evidence (cuobjdump -sass of the objects that make up libnerfb200.so, sm_100a): per kernel, the tensor-core / TMEM /\n"
      "bulk-copy / packed-FMA mnemonics.  UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,\n"
      "HMMA = mma.sync (legacy tensor path), LDSM = ldmatrix, LDGSTS = cp.async, SYNCS = mbarrier ops, FFMA2/FADD2 = packed fp32x2,\n"
      "F2FP = packed float -> half/bf16 conversion, REDG = red.global.add, D* = fp64.\n")
for o in sorted(f for f in os.listdir(OBJ) if f.endswith(".o")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, o)], capture_output=True, text=True).stdout
    print("== " + o)
    name, counts, n = None, collections.Counter(), 0

    def flush():
        if name is None:
            return
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem)
        ks = ", ".join("%s %d" % (k, counts[k]) for k in KEYS if counts[k])
        print("  %-68s %5d instr  %s" % (dem[:68], n, ks))
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, counts, n = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            n += 1
            counts[m.group(1)] += 1
    flush()
