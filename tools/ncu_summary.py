#!/usr/bin/env python
"""Markdown summary of an .ncu-rep (run in the build container: `ncu -i` needs no GPU).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]   > profiles/rNN_xxx.md

Prints, per captured launch, the counters the roofline arguments use (duration, registers, DRAM bytes and throughput,
tensor / FMA / ALU / XU / FP64 pipe utilisation, issue slots, warp occupancy, shared-memory wavefronts, instruction
count, the main stall reasons per issue) and, with --source N, the N source lines with the most executed instructions
(needs -lineinfo at compile time and --import-source on at capture time)."""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nsrc = int(sys.argv[sys.argv.index("--source") + 1]) if "--source" in sys.argv else 0
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")][:60] for r in data]
    print("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    print("| kernel | | " + " | ".join(names) + " |")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("| %s | %s | %s |" % (w, units[i], " | ".join(r[i] for r in data)))
    if not nsrc:
        return
    out = ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"])
    agg, cur, h = collections.OrderedDict(), None, None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            h = r
            ii, si = h.index("Instructions Executed"), h.index("# Samples")
        elif h is not None and r[0].strip().isdigit() and len(r) > ii:
            try:
                n, s = int(r[ii]), int(r[si])
            except ValueError:
                continue
            a = agg.setdefault((cur, int(r[0]), r[1].strip()[:100]), [0, 0])
            a[0] += n
            a[1] += s
    tot = sum(v[0] for v in agg.values()) or 1
    samp = sum(v[1] for v in agg.values()) or 1
    print("\nSource lines by executed warp instructions (all captured launches; %% of %d instructions, %% of %d stall samples):\n" % (tot, samp))
    print("| instr % | samples % | line | source |")
    print("|---|---|---|---|")
    for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:nsrc]:
        print("| %.1f | %.1f | %s:%d | `%s` |" % (100 * n / tot, 100 * s / samp, f, ln, src.replace("|", "\\|")))


if __name__ == "__main__":
    main()
