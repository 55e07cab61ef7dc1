"""Opcode counts per kernel of libosrl_b200.so (cuobjdump -sass): the evidence table profiles/r0X_sass_opcodes.txt.
UTCHMMA = tcgen05.mma (kind::tf32 / f16), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, LDGSTS = cp.async,
UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, HMMA.1688.F32.TF32 = mma.sync m16n8k8 tf32."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "osrl_b200", "libosrl_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UBLKCP", "LDGSTS", "HMMA.1688.F32.TF32", "FFMA", "SYNCS",
        "FENCE.VIEW.ASYNC", "LD.E", "ST.E"]
fn, table = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn).replace("void ", "")
        table[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        op = m.group(1)
        table[fn]["instructions"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                table[fn][c] += 1
print("kernel | instructions | " + " | ".join(cols))
for fn, c in table.items():
    print(f"{fn:70s} | {c['instructions']} | " + " | ".join(str(c[k]) for k in cols))
