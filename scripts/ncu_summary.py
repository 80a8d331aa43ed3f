"""Turn an .ncu-rep into the short text summary committed under profiles/ (run where ncu is installed; no GPU needed)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        print("kernel:", d.get("Kernel Name"), "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for i, name in enumerate(head):
            tensor = "tensor" in name and "pct" in name and ".avg." in name and r[i] not in ("0", "0.000000", "")
            if name in WANT or tensor:
                print("  {0:<75s} {1:>18s} {2}".format(name, r[i], units[i]))
        stalls = sorted(((float(r[i].replace(",", "") or 0), head[i]) for i in range(len(head))
                         if "issue_stalled" in head[i] and head[i].endswith("per_issue_active.ratio")), reverse=True)[:6]
        for v, name in stalls:
            print("  stall {0:<69s} {1:>18.3f}".format(name.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    main(sys.argv[1])
