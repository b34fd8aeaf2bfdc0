"""Extract the metrics the roofline discussion uses from an .ncu-rep (read on the CPU box):
   python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {rep}: selected ncu --set full metrics (per launch)"]
    for r in rows[2:]:
        lines.append("---")
        for w in WANT:
            if w in idx:
                lines.append(f"{w} = {r[idx[w]]} {units[idx[w]]}")
        extra = [h for h in hdr if ("tensor" in h and "pct_of_peak_sustained_elapsed" in h and h not in WANT)]
        for h in extra[:6]:
            lines.append(f"{h} = {r[idx[h]]} {units[idx[h]]}")
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
