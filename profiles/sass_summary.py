"""Per-kernel SASS opcode summary of the built sm_100a library: `python profiles/sass_summary.py > profiles/r2_sass.md`
(cuobjdump -sass bulletproofs_b200/libbpmsm.so).  The columns are the mnemonics that matter for this path: IMAD.WIDE (the 32x32->64
multiply every field product is made of), UBLKCP (cp.async.bulk, the TMA bulk copy) with its mbarrier SYNCS, shuffles, atomics."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bulletproofs_b200", "libbpmsm.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
arch = re.search(r"arch = (sm_\w+)", out)
demangle = subprocess.run(["c++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS summary of `{os.path.relpath(lib, ROOT)}` ({arch.group(1) if arch else '?'}), `cuobjdump -sass`\n")
print("| kernel | instructions | IMAD.WIDE* | other IMAD* | IADD3* | SHF | UBLKCP (TMA bulk) | SYNCS (mbarrier) | SHFL | ATOM/RED | LDG/LD | STG/ST | LDS/STS | LDL/STL |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = collections.Counter()
for (name, c), dn in zip(kern.items(), demangle):
    short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "")).replace("void ", "")
    g = lambda pred: sum(v for k, v in c.items() if pred(k))
    row = [sum(c.values()), g(lambda k: k.startswith("IMAD.WIDE")), g(lambda k: k.startswith("IMAD") and not k.startswith("IMAD.WIDE")), g(lambda k: k.startswith("IADD3")),
           g(lambda k: k.startswith("SHF")), g(lambda k: k.startswith("UBLKCP")), g(lambda k: k.startswith("SYNCS")), g(lambda k: k.startswith("SHFL")),
           g(lambda k: k.startswith("ATOM") or k.startswith("RED")), g(lambda k: k.startswith("LDG") or k.startswith("LD.")), g(lambda k: k.startswith("STG") or k.startswith("ST.")),
           g(lambda k: k.startswith("LDS") or k.startswith("STS")), g(lambda k: k.startswith("LDL") or k.startswith("STL"))]
    print(f"| {short} | " + " | ".join(str(x) for x in row) + " |")
    for k, v in c.items():
        tot[k] += v
print(f"\nWhole library: {sum(tot.values())} instructions, {sum(v for k, v in tot.items() if k.startswith('IMAD.WIDE'))} IMAD.WIDE, "
      f"{sum(v for k, v in tot.items() if k.startswith('UBLKCP'))} UBLKCP, {sum(v for k, v in tot.items() if k.startswith('UTMALDG'))} UTMALDG (tensor-map TMA: none, the copies are 1-D byte ranges), "
      f"no tensor-core instruction (HMMA/IMMA/UTC*: {sum(v for k, v in tot.items() if k.startswith(('HMMA', 'IMMA', 'UTC')))}) -- modular integer arithmetic, not a dense contraction.")
