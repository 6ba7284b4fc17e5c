"""Per-source-line cost of one kernel from an ncu report + the built library's line info.
   python scripts/ncu_lines.py report.ncu-rep mangled_kernel_name [top_n] [per_unit_divisor]
Joins ncu's per-SASS-instruction counters (executed warp instructions, stall samples) with nvdisasm -g line
info of kompressor_b200/libkzg.so and prints the heaviest source lines (inlined code is attributed to the
innermost line)."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "kompressor_b200", "libkzg.so")], cwd=tmp, capture_output=True)
dis = ""
for f in sorted(os.listdir(tmp)):                 # the cubin that holds the kernel
    if f.endswith(".cubin"):
        d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if ".text." + kern in d:
            dis = d
            break
line_of = {}
cur, infn = None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        infn = (".text." + kern + " ") in ln or ln.rstrip("- ").endswith(".text." + kern)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
base_name = re.split(r"IL|Ev8KzgBatch|8KzgBatch|11KzgEncBatch", re.sub(r"^_Z\d+", "", kern))[0]      # ncu matches the un-templated name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", base_name, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[ia], 16), int(r[isamp] or 0), int(r[iex] or 0)) for r in rows[hi + 1:] if len(r) >= len(hdr) and r[ia].startswith("0x")]
base = data[0][0]
agg = collections.defaultdict(lambda: [0, 0, 0])
for a, s, e in data:
    k = line_of.get(a - base, ("?", 0))
    agg[k][0] += s; agg[k][1] += e; agg[k][2] += 1
ts = sum(v[0] for v in agg.values()) or 1
te = sum(v[1] for v in agg.values()) or 1
print(f"total samples {ts}, executed warp instructions {te} ({te / div:.1f} per unit)")
src_cache = {}
def src(f, l):
    if f not in src_cache:
        p = os.path.join(root, "kompressor_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[f]
    return L[l - 1].strip()[:80] if 0 < l <= len(L) else ""
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"  {k[0]}:{k[1]:<4d} exec {v[1] / te * 100:5.1f}% ({v[1] / div:7.1f}/unit) samples {v[0] / ts * 100:5.1f}% sass {v[2]:3d} | {src(*k)}")
