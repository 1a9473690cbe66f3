"""SASS evidence: per-kernel counts of the Blackwell-native mnemonics in randnla_b200/librnla.so (cuobjdump -sass), written to
profiles/r02_sass_summary.txt.  UTCIMMA = tcgen05.mma kind::i8 (.2CTA: cta_group::2), UTMALDG = cp.async.bulk.tensor (tiled TMA), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine),
SYNCS = mbarrier, STAS = st.async (remote shared-memory store completing on the receiver's mbarrier), DMMA = mma.sync f64."""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", "randnla_b200/librnla.so"], capture_output=True, text=True).stdout
pat = ["UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "UTMAPF", "SYNCS", "STAS", "DMMA", "LDGSTS", "REDUX", "ATOMS"]
cur = None
per = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    if cur:
        for p in pat:
            if re.search(r"\b" + p + r"[\. ]", line):
                per[cur][p] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
rows = []
for (k, c), nm in zip(per.items(), names):
    tot.update(c)
    if any(c[p] for p in ("UTCIMMA", "LDTM", "UBLKCP", "UTMALDG", "DMMA")):
        short = nm.replace("rnla::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\((?!anonymous).*", "", short)
        rows.append((short, c))
with open("profiles/r02_sass_summary.txt", "w") as f:
    f.write("cuobjdump -sass randnla_b200/librnla.so (sm_100a), mnemonic counts\n")
    f.write("total over %d kernels: " % len(per) + ", ".join(f"{p} {tot[p]}" for p in pat) + "\n\n")
    f.write(f"{'kernel':<70} " + " ".join(f"{p:>8}" for p in pat) + "\n")
    for short, c in rows:
        f.write(f"{short[:70]:<70} " + " ".join(f"{c[p]:>8}" for p in pat) + "\n")
print(open("profiles/r02_sass_summary.txt").read()[:3500])
