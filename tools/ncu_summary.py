"""Condenses an .ncu-rep (ncu --set full --import-source on) into the text summary kept
under profiles/:  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg ", "launch__registers_per_thread ", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
    "smsp__average_warps_issue_stalled",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path):
    raw = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print("=" * 100)
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for h, u, v in zip(hdr, units, row):
            if any(k.strip() in h for k in KEYS):
                try:
                    if float(v) == 0:
                        continue
                except ValueError:
                    continue
                print(f"  {h:95s} {v:>18s} {u}")
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    if len(src) > 2:
        hdr, data = src[1], src[2:]
        ci = {h: i for i, h in enumerate(hdr)}
        # a multi-kernel report repeats title + header lines per kernel: keep the rows of the LAST kernel
        data = [r for r in data if len(r) >= len(hdr)]
        last = max([i for i, r in enumerate(data) if r[ci["# Samples"]] == "# Samples"], default=-1)
        data = data[last + 1:]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
        tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
        inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
        print("=" * 100)
        print(f"source page of the last kernel: {tot} samples, {inst} warp instructions executed")
        agg = {k: sum(int(r[ci[k]] or 0) for r in data) for k in stalls}
        print("stall totals:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
        print("hottest SASS lines (samples, executions, instruction, top stall reasons):")
        for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:25]:
            st = sorted(((k[6:], int(r[ci[k]] or 0)) for k in stalls), key=lambda x: -x[1])[:2]
            print(f"  {r[ci['# Samples']]:>8s} {r[ci['Instructions Executed']]:>12s}  {r[ci['Source']][:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1])
