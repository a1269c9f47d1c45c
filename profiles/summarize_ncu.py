"""Writes the judged summary of an ncu report: `python profiles/summarize_ncu.py gpurun_out/X.ncu-rep > profiles/X.summary.txt`.
Key counters of every profiled launch (ncu --page raw) plus the hottest source lines (ncu --page source)."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "smsp__sass_inst_executed_op_tmem_stt.sum", "lts__t_bytes.sum"]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised replays: compare shares, not absolutes)")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"\n## {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(d[k] or 0) >= 0.05:
            print(f"{k:75s} {d[k]:>16s}")
