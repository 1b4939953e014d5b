"""Aggregate `ncu --page source --csv --print-source cuda,sass` by source line.
usage: python tools/ncu_lines.py src.csv <units> [top]
units = number of rollout-steps (or whatever) the per-unit figures are divided by."""
import csv
import sys
from collections import defaultdict

path, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file = None
hdr = None
inst = defaultdict(float)
samp = defaultdict(float)
src = {}
stall = defaultdict(lambda: defaultdict(float))
cur_line = None
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "File Name":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(row) < len(hdr):
        continue
    if row[0] != "":
        cur_line = (cur_file, int(row[0]))
        src[cur_line] = row[1]
        continue
    # sass row belonging to cur_line
    try:
        n = float(row[ix["Instructions Executed"]] or 0)
        s = float(row[ix["# Samples"]] or 0)
    except ValueError:
        continue
    inst[cur_line] += n
    samp[cur_line] += s
    for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_lg", "stall_branch_resolving", "stall_math", "stall_no_inst", "stall_mio"):
        try:
            stall[cur_line][k] += float(row[ix[k]] or 0)
        except (ValueError, KeyError):
            pass
tot_i, tot_s = sum(inst.values()), sum(samp.values())
print("total warp instructions %.4g = %.1f per unit; samples %.0f" % (tot_i, tot_i / units, tot_s))
tot_st = defaultdict(float)
for l in stall:
    for k, v in stall[l].items():
        tot_st[k] += v
print("stall samples:", {k: round(v / tot_s, 3) for k, v in tot_st.items()})
print("-- by samples")
for l in sorted(samp, key=lambda l: -samp[l])[:top]:
    st = stall[l]
    big = max(st, key=st.get) if st else ""
    print("%5.1f%% smp %5.1f%% ins %7.1f/unit  %s:%d  [%s]  %s" % (100 * samp[l] / tot_s, 100 * inst[l] / tot_i, inst[l] / units, l[0], l[1], big, src[l].strip()[:110]))
