"""Condensed view of an ncu --set full report: python tools/ncu_summary.py report.ncu-rep [kernel-regex]  (prints selected raw metrics per kernel)."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[0]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed"]
STALL = [n for n in h if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(h, r)); name = d["Kernel Name"]
    if pat and not pat.search(name): continue
    print("==", name[:110])
    for w in WANT:
        if w in d and d[w] != "": print(f"   {w:90s} {d[w]}")
    st = sorted(((float(d[n].replace(',', '')), n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for n in STALL if d.get(n)), reverse=True)[:6]
    print("   top stalls (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in st))
