"""Count the SASS mnemonics that show which hardware paths each kernel uses (B200_PROFILING.md: tcgen05 / TMA evidence).

    cuobjdump -sass str2str_b200/libstr2str_b200.so | python tools/sass_mnemonics.py > profiles/<round>_sass_mnemonics.txt
"""
import collections
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UBLKCP", "UTMASTG", "LDTM", "STTM", "SYNCS", "HMMA", "UTCCP", "ELECT")
cur, counts = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and m.group(1).split(".")[0] in KEYS:
        counts[cur][m.group(1).split(".")[0]] += 1
print("SASS mnemonic counts per kernel of libstr2str_b200.so (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit,")
print("LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = bulk copy, SYNCS = mbarrier ops, UTCATOMSWS = TMEM alloc / dealloc.\n")
for k in sorted(counts, key=lambda k: -counts[k]["UTCHMMA"]):
    c = counts[k]
    if not (c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"]):
        continue
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = name.replace("s2s::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    print(f"{name[:64]:64s} " + " ".join(f"{op}={c[op]}" for op in KEYS if c[op]))
