"""Dev tool: static SASS evidence per kernel of libpartmanip_b200.so — counts of the Blackwell-specific mnemonics
(B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR =
tcgen05.commit, SYNCS = mbarrier ops, UCGABAR = cluster barrier, STAS = st.async (DSMEM push), (C)REDUX = warp reductions.

    python scripts/sass_evidence.py > profiles/<name>.txt        (needs cuobjdump + c++filt; no GPU)
"""
import collections
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "partmanip_b200", "csrc", "libpartmanip_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|STAS|CREDUX|REDUX|MUFU\.TANH|HMMA)\b")
counts, order, fn = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        order.append(fn)
        continue
    if fn:
        for k in pat.findall(line):
            counts[fn][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass partmanip_b200/csrc/libpartmanip_b200.so (sm_100a): Blackwell-specific mnemonics per kernel")
print("# (template instances of one kernel with identical counts are merged)")
seen = collections.OrderedDict()
for f, n in zip(order, names):
    c = counts[f]
    if not any(k.startswith("UTC") or k in ("LDTM", "STAS", "UCGABAR_ARV", "SYNCS") for k in c):
        continue
    m = re.search(r"([A-Za-z_0-9]+)(<[^(]*>)?\(", n.replace("(anonymous namespace)::", ""))
    base = m.group(1) if m else n
    key = (base, tuple(sorted(c.items())))
    seen[key] = seen.get(key, 0) + 1
for (base, items), k in seen.items():
    print(f"{base:34s} x{k:<3d} " + "  ".join(f"{a}={b}" for a, b in items))
