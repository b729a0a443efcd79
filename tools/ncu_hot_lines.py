"""Per-source-line stall sampling summary from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_hot_lines.py report.ncu-rep <kernel-regex> [top]"""
import csv, io, re, subprocess, sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--kernel-name", f"regex:{kre}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
# the CSV holds blocks: file header lines, then a table with per-line metrics
lines = out.splitlines()
hdr_i = [i for i, l in enumerate(lines) if l.startswith('"Line No","Source"') or l.startswith('"#","Source"') or ('"Source"' in l and 'Sampl' in l)]
blocks = []
cur_file = None
for i, l in enumerate(lines):
    if l.startswith('"File Name"'):
        cur_file = l.split(",", 1)[1].strip('"')
    if '"Source"' in l and "Samp" in l:
        hdr = next(csv.reader([l]))
        j = i + 1
        rows = []
        while j < len(lines) and not lines[j].startswith('"File Name"') and lines[j].strip():
            try:
                rows.append(next(csv.reader([lines[j]])))
            except Exception:
                pass
            j += 1
        blocks.append((cur_file, hdr, rows))
if not blocks:
    print(out[:2000]); sys.exit(0)
for f, hdr, rows in blocks:
    si = [k for k, h in enumerate(hdr) if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)" or "Sampling (All" in h]
    if not si:
        print("columns:", hdr[:40]); continue
    si = si[0]
    tot = sum(float(r[si] or 0) for r in rows if len(r) > si)
    if tot == 0: continue
    print(f"== {f}  total samples {tot:.0f}")
    best = sorted(rows, key=lambda r: -float(r[si] or 0))[:top]
    for r in best:
        print(f"{float(r[si]) / tot:6.1%}  L{r[0]:>4s}  {r[1][:130]}")
