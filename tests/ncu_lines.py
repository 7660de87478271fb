"""Diagnostics: stall samples per CUDA source line from a .ncu-rep captured with --import-source on (top N lines).
   python tests/ncu_lines.py report.ncu-rep [N] [file-substring]"""
import csv, subprocess, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
only = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
lines, cur = {}, ""
for r in rows:
  if len(r) >= 2 and r[0] == "File Path":
    cur = r[1].split("/")[-1]
  if len(r) == len(hdr) and r[0].isdigit() and r[si].replace(".", "").isdigit():
    k = (cur, int(r[0])); s, n = lines.get(k, (0, 0, ""))[:2]
    lines[k] = (s + float(r[si]), n + float(r[ii] or 0), r[1])
tot = sum(v[0] for v in lines.values())
print("total samples", tot)
sel = {k: v for k, v in lines.items() if only in k[0]}
print("selected", sum(v[0] for v in sel.values()))
for k, (s, n, src) in sorted(sel.items(), key=lambda kv: -kv[1][0])[:top]:
  print("%-18s %5d %6.1f%% %9d inst  %s" % (k[0][:18], k[1], 100 * s / tot, n, src.strip()[:100]))
