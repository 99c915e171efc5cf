"""Print the metrics the roofline / VERDICT discussion uses from an ncu report (run where ncu is installed, no GPU needed):
    python scripts/ncu_extract.py gpurun_out/r02_gemm2cta.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"== {path}")
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:<72} {r[i][:100]} {units[i]}")
        print()
