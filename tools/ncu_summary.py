"""Key metrics of every kernel in an .ncu-rep as a small markdown table (committed under profiles/)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__cycles_elapsed.max", "cycles")]
idx = [(hdr.index(m), lab) for m, lab in want if m in hdr]
print("| " + " | ".join(l for _, l in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, lab in idx:
        v = r[i]
        if lab == "kernel":
            v = v.split("(")[0][-40:]
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g} {units[i]}".strip()
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
