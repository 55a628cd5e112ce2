"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the small JSON files kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r1_ncu_x.json "note text" """
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_selected"]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    kernels = []
    for row in rows[2:]:
        k = {"Kernel Name": row[head.index("Kernel Name")]}
        for key in KEYS:
            if key in head:
                k[key] = row[head.index(key)] + " " + units[head.index(key)]
        kernels.append(k)
    json.dump({"note": note, "source": rep, "kernels": kernels}, open(out, "w"), indent=1)
    print(f"wrote {out}: {len(kernels)} kernel(s)")


if __name__ == "__main__":
    main()
