"""Per-source-line instruction and stall-sample totals from an .ncu-rep (needs -lineinfo builds)."""
import csv
import subprocess
import sys
import io

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) >= 8 and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0)))
        except ValueError:
            pass
tot_i = sum(l[4] for l in lines) or 1
tot_s = sum(l[3] for l in lines) or 1
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
print("%-18s %5s %7s %7s  %s" % ("file", "line", "inst%", "samp%", "source"))
for f, ln, src, smp, ins in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%-18s %5d %6.2f%% %6.2f%%  %s" % (f, ln, 100.0 * ins / tot_i, 100.0 * smp / tot_s, src[:110]))
