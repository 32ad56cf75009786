"""Summarise an .ncu-rep (read here, no GPU) into the few numbers DESIGN.md / bench.py quote."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                print(f"  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
        stalls = sorted(((float(r[i]), h) for i, h in enumerate(hdr)
                         if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[i]), reverse=True)[:6]
        print("  top stall reasons (samples):", ", ".join(f"{h.split('stalled_')[1]}={int(v)}" for v, h in stalls))


if __name__ == "__main__":
    main(sys.argv[1])
