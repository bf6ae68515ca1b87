#!/usr/bin/env python
"""Compact text summary of an Nsight Compute report (run here, no GPU needed).

    python tools/ncu_report.py gpurun_out/prof.ncu-rep [kernel-substring] [--top 14]

Prints, per profiled kernel: duration, DRAM bytes, throughput percentages, occupancy, the top warp
stall reasons, and the hottest CUDA source lines (needs -lineinfo + --import-source on).
The judged copies live under profiles/.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 ld sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 red sectors"),
    ("lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "L2 atomic unit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        if filt and filt not in name:
            continue
        if name in seen:
            continue
        seen.add(name)
        print("=" * 100)
        print("kernel", d["Kernel Name"][:90], " id", d.get("ID"))
        for k, label in KEYS:
            if k in d:
                print("  %-22s %s %s" % (label, d[k], units[hdr.index(k)]))
        st = []
        for k in hdr:
            if "issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[k].replace(",", "")), k.split("issue_stalled_")[1].split("_per_")[0]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
        src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + name.split("::")[-1].split("<")[0].split()[-1]])
        lines, cur_file = [], ""
        H = None
        for rr in csv.reader(io.StringIO(src)):
            if not rr:
                continue
            if rr[0] == "File Path":
                cur_file = rr[1].split("/")[-1]
            elif rr[0] == "Line No":
                H = rr
            elif H and rr[0].isdigit():
                dd = dict(zip(H, rr))
                try:
                    lines.append((int(dd["# Samples"]), int(dd["Instructions Executed"]), cur_file, rr[0], rr[1].strip()[:95]))
                except (ValueError, KeyError):
                    pass
        ts = sum(l[0] for l in lines) or 1
        ti = sum(l[1] for l in lines) or 1
        print("  hottest source lines (%% stall samples, %% warp instructions):")
        for smp, ins, f, ln, text in sorted(lines, reverse=True)[:top]:
            print("   %5.1f%% %5.1f%%  %s:%s  %s" % (100. * smp / ts, 100. * ins / ti, f, ln, text))


if __name__ == "__main__":
    main()
