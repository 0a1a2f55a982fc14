"""Key metrics of an `ncu --set full` report, one block per captured launch:
    python tools/ncu_key_metrics.py gpurun_out/r2_iteration_full.ncu-rep > profiles/r2_iteration_kernels_ncu_full_key_metrics.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_sample_buffers" ]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("==== %s  (launch id %s)" % (r[idx["Kernel Name"]], r[idx["ID"]]))
    for w in WANT:
        if w in idx and r[idx[w]] != "":
            print("  %-72s %16s %s" % (w, r[idx[w]], units[idx[w]]))
    rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
    ur, uw = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
    sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tot = rd * sc[ur] + wr * sc[uw]
    dur = float(r[idx["gpu__time_duration.sum"]]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[idx["gpu__time_duration.sum"]]]
    print("  %-72s %16.4f GB  -> %.0f GB/s over the launch" % ("dram read + write", tot / 1e9, tot / dur / 1e9))
