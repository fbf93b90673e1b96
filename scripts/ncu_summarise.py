"""Development aid: pick the judged metrics out of `ncu -i X.ncu-rep --page raw --csv` output.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > profiles/rNN_X_raw.csv
    python scripts/ncu_summarise.py profiles/rNN_X_raw.csv
"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_ncu_peak",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
}


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT and vals[i] != "":
                d[WANT[h]] = "%s %s" % (vals[i], units[i])
        out.append(d)
    return out


if __name__ == "__main__":
    print(json.dumps(summarise(sys.argv[1]), indent=1))
