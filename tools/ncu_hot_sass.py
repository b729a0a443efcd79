"""Stall-sampling summary per SASS instruction from an .ncu-rep.
usage: python tools/ncu_hot_sass.py report.ncu-rep <kernel-regex> [top]"""
import csv, subprocess, sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel block only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
si = hdr.index("# Samples")
ie = hdr.index("Instructions Executed")
stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[si] or 0) for r in body)
print(f"{len(body)} SASS instructions, {tot:.0f} samples, {sum(float(r[ie] or 0) for r in body):.0f} warp-instructions executed")
agg = {}
for r in body:
    for k, h in stall_cols:
        agg[h] = agg.get(h, 0.0) + float(r[k] or 0)
print("stall mix:", ", ".join(f"{h[6:]}={v / tot:.1%}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -float(body[i][si] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((float(r[k] or 0), h[6:]) for k, h in stall_cols), reverse=True)[:2]
    print(f"{float(r[si]) / tot:6.1%} #{i:5d} exec={float(r[ie] or 0):9.0f}  {r[1].strip()[:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
