"""Selected columns of an `ncu --set full` report as a small transposed CSV (metric, unit, value per kernel) for profiles/.
    python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration|dram__bytes_(read|write)\.sum|dram__throughput|lts__t_bytes\.sum($|\.per_second)|lts__t_sector_hit_rate|"
                  r"launch__|sm__throughput|sm__inst_executed\.sum($|\.per_cycle)|sm__warps_active|sm__cycles_active\.avg|"
                  r"sm__pipe_tensor|sm__inst_executed_pipe_(tensor|uniform|lsu|alu|fma)|sm__mem_tensor|smsp__inst_executed\.sum$|"
                  r"l1tex__data_bank_conflicts|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|l1tex__t_bytes\.sum$|"
                  r"smsp__average_warp.*issue_stalled.*(long_scoreboard|barrier|membar|short_scoreboard|wait|sleeping|math_pipe|lg_throttle)|"
                  r"smsp__cycles_active\.avg|sm__sass_inst_executed_op_shared|smsp__warp_issue_stalled.*_per_warp_active)")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[name_i][:60] for r in data])
        for i, h in enumerate(hdr):
            if KEEP.search(h):
                w.writerow([h, units[i]] + [r[i] for r in data])
    print(out, sum(1 for _ in open(out)), "rows")


if __name__ == "__main__":
    main()
