"""SASS opcode histogram of the hot kernels in kompressor_b200/libkzg.so (cuobjdump -sass), for profiles/.
   python scripts/sass_hist.py [kernel-name-substring ...] > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "kompressor_b200", "libkzg.so")
want = sys.argv[1:] or ["kzg_exec_kernel", "kzg_exec4_kernel", "kzg_fse_kernel", "kzg_huf_kernel", "kzg_fsetab_kernel", "kzg_parse_kernel",
                        "kzg_bulk_kernel", "kzg_finish_kernel", "kzg_xxh_segment_kernel"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, {}
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        hist.setdefault(cur, collections.Counter())[m.group(1)] += 1
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
notable = ("LDGSTS", "UBLKCP", "UTMA", "LDGDEPBAR", "DEPBAR", "ATOMS", "ATOMG", "RED", "REDUX", "BAR", "MEMBAR", "SHFL", "VOTE", "MATCH", "LDS", "STS",
           "LDG", "STG", "SHF", "LOP3", "PRMT", "BSSY", "BRA", "NANOSLEEP", "FENCE", "SYNCS")
for k in sorted(hist):
    if not any(w in k for w in want):
        continue
    h = hist[k]
    tot = sum(h.values())
    base = collections.Counter()
    for op, c in h.items():
        base[op.split(".")[0]] += c
    print(f"{demangle(k)}\n  {tot} instructions; by opcode: " + ", ".join(f"{op} {c}" for op, c in base.most_common(22)))
    print("  notable: " + ", ".join(f"{op} {sum(c for o, c in h.items() if o.split('.')[0].startswith(op))}" for op in notable
                                     if any(o.split('.')[0].startswith(op) for o in h)))
    full = [f"{o} {c}" for o, c in h.most_common() if o.startswith(("LDGSTS", "UBLKCP", "ATOMS", "REDUX", "BAR", "MEMBAR", "LDGDEPBAR", "DEPBAR"))]
    if full:
        print("  async / sync forms: " + ", ".join(full))
    print()
