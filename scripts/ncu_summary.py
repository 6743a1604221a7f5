#!/usr/bin/env python
"""Text summary of an .ncu-rep (the files under profiles/): a fixed list of raw metrics per kernel, DRAM traffic per launch,
and the hottest SASS instructions by warp-stall samples.      python scripts/ncu_summary.py report.ncu-rep > summary.txt
Reads the report with `ncu -i ... --page raw --csv` / `--page source --csv` (the recipe of B200_PROFILING.md)."""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio")
PREFIX = ("smsp__average_warps_issue_stalled_", "dram__bytes_read.sum.", "dram__bytes_write.sum.")


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return [r for r in csv.reader(io.StringIO(out))]


def main():
    rep = sys.argv[1]
    print(f"# ncu --set full --clock-control none --import-source on  ({rep.split('/')[-1]})")
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"## kernel: {r[kn]}")
        vals = {}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP or h.startswith(PREFIX):
                print(f"{h:<95} {u:<12} {v}")
                vals[h] = (u, v)
        try:
            def gb(k):
                u, v = vals[k]
                return float(v) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}[u]
            print(f"traffic (dram read + write) per launch: {gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum'):.4f} GB")
        except (KeyError, ValueError):
            pass
    src = ncu(rep, "source")
    # one block per kernel: a "Kernel Name" line, a header line, then instructions
    i = 0
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            name = src[i][1]
            h = src[i + 1]
            j = i + 2
            body = []
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                body.append(src[j])
                j += 1
            try:
                isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
                stall_cols = [(k, c) for k, c in enumerate(h) if c.startswith("stall_")]
                data = [(int(b[isamp] or 0), b) for b in body if len(b) > isamp and b[isamp].isdigit()]
                tot = sum(d[0] for d in data) or 1
                print(f"\n# hottest SASS instructions by warp-stall samples (ncu --page source): {name[:80]}")
                print(f"total samples {tot}, {len(data)} SASS instructions")
                for s, b in sorted(data, key=lambda d: -d[0])[:16]:
                    top = max(stall_cols, key=lambda kc: int(b[kc[0]] or 0) if b[kc[0]].isdigit() else 0)[1] if stall_cols else ""
                    print(f"{100 * s / tot:5.1f}%  {top:<22} exec={b[iex]:>9}  {b[isrc].strip()[:90]}")
            except ValueError:
                pass
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main()
