"""Tensor-core / TMA opcode counts per kernel of libbvlm.so (evidence that the hot path is tcgen05 + TMA code):
    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
Counts SASS mnemonics from `cuobjdump -sass`: UTCHMMA / UTCQMMA (tcgen05.mma kind::f16 / kind::f8f6f4), LDTM (tcgen05.ld),
UTMALDG / UTMASTG (cp.async.bulk.tensor load / store), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
MUFU, plus legacy HMMA (must be zero: no mma.sync fallback)."""
import collections, re, subprocess, sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "bayesvlm_b200" / "libbvlm.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU", "HMMA", "LDGSTS"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k) and not (k == "HMMA" and op.startswith("UTCHMMA")):
                per[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib.name}: {len(per)} kernels; columns: total instructions, then {' '.join(KEYS)}")
tot = collections.Counter()
for (mangled, c), name in zip(per.items(), names):
    tot.update(c)
    short = re.sub(r"\(anonymous namespace\)::|bvlm::", "", name)
    short = re.sub(r"\(CUtensorMap_st.*", "", short)[:110]
    print(f"{short:112s} {c['_total']:6d} " + " ".join(f"{c[k]:5d}" for k in KEYS))
print(f"{'TOTAL':112s} {tot['_total']:6d} " + " ".join(f"{tot[k]:5d}" for k in KEYS))
