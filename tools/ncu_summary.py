"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (ncu --set full)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_hmma%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs")]
cols = [(hdr.index(k), n) for k, n in want if k in hdr]
units = rows[1]
for r in rows[2:]:
    print(" | ".join("%s=%s%s" % (n, r[i][:44] if n == "kernel" else r[i], (" " + units[i]) if n in ("us", "dram_rd", "dram_wr", "l2_bytes") else "") for i, n in cols))
if len(sys.argv) > 2:
    for i, h in enumerate(hdr):
        if sys.argv[2] in h:
            print(h, units[i], [r[i] for r in rows[2:]])
