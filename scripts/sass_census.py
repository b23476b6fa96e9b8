"""SASS census of dr-nmf_b200/libdrnmf.so: Blackwell-native instruction counts per kernel (cuobjdump -sass).
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk."""
import collections, os, re, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "dr-nmf_b200", "libdrnmf.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "UCGABAR", "UTMACCTL"]
tot = collections.Counter(); per = collections.OrderedDict(); cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    if cur is None: continue
    for o in ops:
        if re.search(r"\b" + re.escape(o) + r"\b", line):
            tot[o] += 1; per[cur][o] += 1
dem = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
print("# SASS census of dr-nmf_b200/libdrnmf.so (`python scripts/sass_census.py`, final round-2 build)\n")
print("```")
for o in ops: print("%-16s %6d" % (o, tot[o]))
print("```\n\nPer kernel (kernels without any of these instructions are omitted):\n```")
for (k, c), d in zip(per.items(), dem):
    if not sum(c.values()): continue
    name = re.sub(r"\(.*", "", d).replace("void ", "")
    print("%-60s %s" % (name[:60], "  ".join("%s %d" % (o, c[o]) for o in ops[:5])))
print("```")
