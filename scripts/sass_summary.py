"""Per-kernel SASS mnemonic counts of libyms_b200.so (cuobjdump -sass): the evidence that the tensor-core kernels use tcgen05
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit), TMA (UTMALDG / UTMASTG) and mbarriers (SYNCS).
Runs without a GPU.  Usage: python scripts/sass_summary.py [r01] -> profiles/sass_<tag>.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
lib = os.path.join(ROOT, "yolo_ms_b200", "libyms_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "MUFU", "SHFL", "VOTE", "LDS", "STS", "LDG", "STG", "FFMA", "NANOSLEEP"]
cur, counts = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        counts[cur]["_n"] += 1
        for k in keys:
            if m.group(1).startswith(k):
                counts[cur][k] += 1
names = list(counts)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
out = [f"# SASS mnemonic counts per kernel (`cuobjdump -sass yolo_ms_b200/libyms_b200.so`, sm_100a; {tag})\n",
       "UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier ops.\n",
       "| kernel | instrs | " + " | ".join(keys) + " |", "|---|---|" + "---|" * len(keys)]
for n, d in sorted(zip(names, dem), key=lambda x: x[1]):
    d = d.replace("(anonymous namespace)::", "").replace("yms::", "").replace("void ", "")
    d = re.sub(r"\(.*$", "", d)
    c = counts[n]
    out.append(f"| `{d}` | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
open(os.path.join(ROOT, "profiles", f"sass_{tag}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
