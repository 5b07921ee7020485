"""Print the roofline-relevant metrics of every launch in an .ncu-rep (read with `ncu -i <rep> --page raw --csv`) as a small table."""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "dur"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_act%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_el%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "ingest"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"), ("smsp__inst_executed.sum", "inst")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in out.split("\n") if l.startswith('"')]))
hdr, units = rows[0], rows[1]
idx = {n: hdr.index(n) for n, _ in WANT if n in hdr}
print("| kernel | " + " | ".join(f"{s} [{units[idx[n]]}]" for n, s in WANT if n in idx) + " |")
for r in rows[2:]:
    print("| `" + r[hdr.index("Kernel Name")][:110] + "` | " + " | ".join(r[idx[n]] for n, _ in WANT if n in idx) + " |")
