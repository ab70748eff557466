#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the
roofline argument needs.  Usage: scripts/ncu_summary.py file.ncu-rep [...]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (legacy ctr)"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tcgen05 pipe % of elapsed"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tcgen05 pipe % of active"),
    ("sm__inst_executed_pipe_tma.sum", "TMA instrs"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts by tensor core %"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed", "smem pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("smsp__inst_executed.sum", "instructions"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print("#", path)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            print("kernel %s  (launch id %s)" % (name, r[hdr.index("ID")]))
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    print("    %-34s %s %s" % (label, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_warp_active.pct", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            if stalls:
                print("    top stalls: " + ", ".join("%s %.0f%%" % (n, v) for v, n in stalls[:4]))


if __name__ == "__main__":
    main()
