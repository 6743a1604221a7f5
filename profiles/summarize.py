"""Turn an .ncu-rep (captured on the B200 box under gpurun) into the text summary committed here.
    python profiles/summarize.py gpurun_out/prof_tc_v2.ncu-rep profiles/r01/fit_tc_kernel.txt
"""
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hot_sass  # noqa: E402

KEYS = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
)


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none --import-source on  ({os.path.basename(rep)})"]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"## kernel: {name}")
        for i, h in enumerate(hdr):
            if any(h == k or h.startswith(k + ".") for k in KEYS) or (h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h):
                lines.append(f"{h:95s} {units[i]:12s} {r[i]}")
        rd = float(r[hdr.index("dram__bytes_read.sum")])
        wr = float(r[hdr.index("dram__bytes_write.sum")])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tr = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
        lines.append(f"traffic (dram read + write) per launch: {tr / 1e9:.4f} GB")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = out + ".src.csv"
    with open(tmp, "w") as f:
        f.write(src)
    lines.append("\n# hottest SASS instructions by warp-stall samples (ncu --page source)")
    buf = io.StringIO()
    so = sys.stdout
    sys.stdout = buf
    try:
        hot_sass.main(tmp, 24)
    finally:
        sys.stdout = so
    os.remove(tmp)
    lines.append(buf.getvalue())
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
