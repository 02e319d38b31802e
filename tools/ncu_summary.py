"""Summarise an .ncu-rep (ncu --set full) into the text kept under profiles/: python tools/ncu_summary.py rep.ncu-rep > out.txt"""
import csv
import subprocess
import sys

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (per launch)")
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("----")
        print("Kernel Name =", d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                print(f"{k} [{u[k]}] = {d[k]}")
        try:
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
            rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
            print(f"traffic (dram read + write) [Mbyte] = {rd + wr:.3f}")
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main()
