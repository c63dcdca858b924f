"""Summarise an ncu report (run where ncu is installed; no GPU needed): key raw metrics, stall
reasons and the hottest source lines of the first captured launch."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_mh.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, r = rows[0], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"]
for k in keys:
    if k in hdr:
        print("%-70s %s %s" % (k, r[hdr.index(k)], rows[1][hdr.index(k)]))
st = []
for i, h in enumerate(hdr):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            st.append((float(r[i]), h.split("stalled_")[1]))
        except ValueError:
            pass
tot = sum(v for v, _ in st) or 1
print("stalls:", ", ".join("%s %.1f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(src.splitlines()))
hidx = [i for i, x in enumerate(rows) if x and x[0] == "Line No"]
agg = collections.OrderedDict()
first_files = set()
for si, hi in enumerate(hidx):
    f = rows[hi - 2][1].split("/")[-1]
    if f in first_files:
        break  # second launch
    first_files.add(f)
    h = rows[hi]
    isamp, iinst = h.index("# Samples"), h.index("Instructions Executed")
    end = hidx[si + 1] - 2 if si + 1 < len(hidx) else len(rows)
    for x in rows[hi + 1:end]:
        if len(x) > iinst and x[0] != "":
            key = (f, int(x[0]), x[1].strip()[:80])
            a = agg.setdefault(key, [0, 0])
            try:
                a[0] += int(x[iinst] or 0)
                a[1] += int(x[isamp] or 0)
            except ValueError:
                pass
ti, ts = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * v[1] / ts, 100 * v[0] / ti, k[0], k[1], k[2]))
