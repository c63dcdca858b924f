"""Opcode histogram (executed warp instructions and stall samples) of one kernel from `ncu --page source --csv`.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python scripts/sass_hist.py src.csv [chain_steps_per_launch]
"""
import collections
import csv
import re
import sys


def main(path, units=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    i_src, i_n, i_s = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    i_thr = hdr.index("Thread Instructions Executed")
    ops, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src].strip())
        op = (m.group(2) if m else r[i_src].strip()).split(".")[0]
        ops[op] += int(r[i_n])
        samp[op] += int(r[i_s])
        thr[op] += int(r[i_thr])
    tot, ts = sum(ops.values()), max(1, sum(samp.values()))
    print("warp instructions %d  thread instructions %d  lanes %.1f" % (tot, sum(thr.values()), sum(thr.values()) / tot))
    if units:
        print("warp instructions per unit %.2f" % (tot / float(units)))
    for k, v in ops.most_common(45):
        print("%-10s %6.2f%% instr  %6.2f%% samples" % (k, 100.0 * v / tot, 100.0 * samp[k] / ts))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
