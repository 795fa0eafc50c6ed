"""Summarise an .ncu-rep: headline metrics per kernel and the hottest source lines.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex] [top-lines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "scs_"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25

WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__block_size", "launch__grid_size"]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
for d in data:
    name = d[hdr.index("Kernel Name")]
    print("====", name[:110])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("  %-85s %s %s" % (w, d[i][:20], units[i]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
cache = {}


def line_of(path, ln):
    if path not in cache:
        try:
            cache[path] = open(path).read().split("\n")
        except OSError:
            cache[path] = None
    s = cache[path]
    return s[ln - 1].strip()[:95] if s and ln - 1 < len(s) else ""


kernels = collections.OrderedDict()
for n, i in enumerate(hidx):
    kname, path, h = rows[i - 1][1], rows[i - 2][1], rows[i]
    iL, iI, iS = h.index("Line No"), h.index("Instructions Executed"), h.index("# Samples")
    end = hidx[n + 1] - 2 if n + 1 < len(hidx) else len(rows)
    agg = kernels.setdefault(kname, collections.defaultdict(lambda: [0, 0]))
    for r in rows[i + 1:end]:
        if len(r) < len(h):
            continue
        try:
            ln, ins, sm = int(r[iL]), int(r[iI] or 0), int(r[iS] or 0)
        except ValueError:
            continue
        agg[(path, ln)][0] += ins
        agg[(path, ln)][1] += sm
for kname, agg in kernels.items():
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print("==== hottest lines:", kname[:90], "instructions", tot, "samples", tots)
    for (path, ln), (ins, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-16s:%4d %5.1f%%inst %5.1f%%samples | %s" % (path.split("/")[-1][:16], ln, 100 * ins / tot, 100 * sm / tots, line_of(path, ln)))
