"""One-line-per-kernel summary of an ncu report (duration, DRAM bytes, throughput, hit rates, occupancy).
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(f"ncu -i {sys.argv[1]} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum"]
for r in rows[2:]:
    print("==", r[ci["Kernel Name"]][:100])
    for w in want:
        if w in ci: print(f"   {w:75s} {r[ci[w]]:>16s} {rows[1][ci[w]]}")
