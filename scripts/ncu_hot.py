"""Hot SASS regions of one kernel: python scripts/ncu_hot.py report.ncu-rep kernel_name [top_n]
Shows the share of stall samples by SASS address range (buckets of 64 instructions) and the top instructions."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[ia].startswith("0x"):
        continue
    data.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), {h: int(r[i] or 0) for i, h in stall_cols}))
if not data:
    sys.exit("no rows")
base = data[0][0]
tot = sum(d[2] for d in data) or 1
print(f"total samples {tot}, instructions {len(data)}")
B = 48
for k in range(0, len(data), B):
    chunk = data[k:k + B]
    s = sum(d[2] for d in chunk)
    ex = max(d[3] for d in chunk)
    if s / tot > 0.01:
        print(f"  [{k:5d}..{k + len(chunk):5d}) off 0x{chunk[0][0] - base:05x}  {s / tot * 100:5.1f}%  max_exec {ex}")
print("top instructions:")
for d in sorted(data, key=lambda x: -x[2])[:top]:
    st = ", ".join(f"{h[6:]} {v}" for h, v in sorted(d[4].items(), key=lambda x: -x[1])[:2] if v)
    print(f"  0x{d[0] - base:05x} {d[2] / tot * 100:5.2f}% ex {d[3]:9d}  {d[1][:70]:70s} {st}")
