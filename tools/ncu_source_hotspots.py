"""Rank the source lines of one profiled launch by warp-stall samples (CPU-side reading of an ncu report):
  python tools/ncu_source_hotspots.py gpurun_out/prof.ncu-rep <launch index> [top N]
needs the kernels compiled with -lineinfo and the capture taken with --import-source on."""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, cur, named = [], None, False
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) == 2 and r[0] == "Function Name":
        if not named:
            print("#", r[1][:160])
        named = True
    elif len(r) > 5 and r[0] not in ("", "Line No"):
        try:
            lines.append((int(r[4]), cur, int(r[0]), r[1].strip()[:120]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
print(f"# {tot} stall samples")
for s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / tot:5.1f}%  {f}:{ln}  {src}")
