"""Per-source-line executed instructions and stall samples from `ncu --page source --print-source cuda,sass --csv`.

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv > lines.csv ; python scripts/line_hist.py lines.csv [top]
"""
import csv
import os
import sys


def main(path, top=40):
    cur, hdr, agg = None, None, []
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1])
            continue
        if r[0] == "Line No":
            hdr = r
            i_n, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if r[0] == "Function Name" or hdr is None or len(r) < len(hdr):
            continue
        if r[2] != "-":  # SASS rows repeat under their source line
            continue
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:90], int(r[i_n]), int(r[i_s])))
        except ValueError:
            pass
    tot, ts = sum(a[3] for a in agg), max(1, sum(a[4] for a in agg))
    print("total warp instructions %d, samples %d" % (tot, ts))
    for f, ln, src, n, s in sorted(agg, key=lambda a: -a[3])[:top]:
        print("%5.2f%% instr %5.2f%% samp  %s:%d  %s" % (100.0 * n / tot, 100.0 * s / ts, f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
