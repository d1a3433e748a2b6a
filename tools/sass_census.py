"""SASS instruction census per kernel of libpixelforge.so (cuobjdump -sass): instructions, registers are in build/pfcu.ptxas.log.
Columns that matter for the profiling recipe: vector global access (LDG/STG .128), async copies (LDGSTS = cp.async, UBLKCP = cp.async.bulk
through the TMA unit, SYNCS = mbarrier), warp primitives (MATCH, VOTE, SHFL, REDUX), packed integer dot products (IDP), conversions (I2F/F2I on XU)."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "pixelforge_b200/lib/libpixelforge.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None; per = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        per[fn]["total"] += 1
        base = op.split(".")[0]
        per[fn][base] += 1
        if base in ("LDG", "STG") and ".128" in op: per[fn][base + ".128"] += 1
cols = ["total", "LDG", "LDG.128", "STG", "STG.128", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "MATCH", "VOTE", "SHFL", "REDUX", "IDP", "I2F", "I2FP", "F2I", "FRND", "FMUL", "FADD", "IMAD", "LOP3", "PRMT"]
print("%-58s" % "kernel" + "".join("%8s" % c for c in cols))
for fn, c in per.items():
    print("%-58s" % fn[:57] + "".join("%8d" % c.get(k, 0) for k in cols))
