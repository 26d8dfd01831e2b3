"""Per-source-line executed-instruction table from an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py report.ncu-rep kernel-regex units [top] [samples]

A fifth argument `samples` sorts by stall samples instead of executed instructions.

`units` = the number of work units (rays) the launch processed, so the table reads "warp
instructions per ray"."""
import csv
import subprocess
import sys

rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by_samples = len(sys.argv) > 5 and sys.argv[5] == "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
agg, stall, cur, hdr = {}, {}, None, None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None:
        continue
    try:
        ln, n, sm = int(r[0]), int(r[ie]), int(r[isamp])
    except (ValueError, IndexError):
        continue
    key = (cur.split("/")[-1], ln, r[1].strip()[:100])
    agg[key] = agg.get(key, 0) + n
    stall[key] = stall.get(key, 0) + sm
tot, stot = sum(agg.values()), max(1, sum(stall.values()))
print("total warp instructions %d = %.1f per unit" % (tot, tot / units))
for k, v in sorted(agg.items(), key=lambda kv: -(stall[kv[0]] if by_samples else kv[1]))[:top]:
    print("%7.1f  %4.1f%% samples  %s:%d  %s" % (v / units, 100.0 * stall[k] / stot, k[0], k[1], k[2]))
