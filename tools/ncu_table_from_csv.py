"""profiles/ncu_*_kernels_r01.csv (tools/ncu_kernel_table.sh) -> profiles/ncu_kernel_table_r01.json: one row per kernel (its last
captured launch) with the counters DESIGN.md section 4 quotes.  python tools/ncu_table_from_csv.py profiles/*.csv > out.json"""
import collections
import csv
import json
import sys

COLS = {"gpu__time_duration.sum": ("us", 1e-3), "dram__bytes_read.sum": ("rdMB", None), "dram__bytes_write.sum": ("wrMB", None),
        "lts__t_sector_hit_rate.pct": ("L2hit%", 1), "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed": ("alu%", 1),
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": ("fmaheavy%", 1),
        "sm__warps_active.avg.pct_of_peak_sustained_active": ("warps%", 1),
        "smsp__issue_active.avg.pct_of_peak_sustained_active": ("issue%", 1),
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct": ("longsb%", 1),
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct": ("noinst%", 1)}
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rows = collections.OrderedDict()
    for path in sys.argv[1:]:
        for r in csv.reader(open(path)):
            if len(r) < 15 or not r[0].isdigit():
                continue
            key = (path, r[0])
            name = r[4].split("(")[0].replace("void ", "").replace("qq::", "")
            e = rows.setdefault(key, {"kernel": name, "grid": r[8], "block": r[7]})
            metric, unit, val = r[12], r[13], float(r[14].replace(",", ""))
            if metric in COLS:
                col, scale = COLS[metric]
                e[col] = round(val * (UNIT.get(unit, 1.0) if scale is None else scale), 1)
    last = collections.OrderedDict()
    for e in rows.values():
        last[(e["kernel"], e["grid"])] = e          # the last launch of each (kernel, grid) wins
    print(json.dumps(list(last.values()), indent=0))


if __name__ == "__main__":
    main()
