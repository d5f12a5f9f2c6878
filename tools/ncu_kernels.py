"""Print the key raw metrics of every kernel in an ncu report: python tools/ncu_kernels.py <report.ncu-rep> [metric substrings...]"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct",
                        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct", "sm__warps_active.avg.pct",
                        "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "smsp__cycles_active.avg",
                        "sm__cycles_elapsed.avg ", "gpc__cycles_elapsed.max", "launch__occupancy_limit", "sm__throughput.avg.pct",
                        "stalled_short_scoreboard_per", "stalled_long_scoreboard_per", "stalled_barrier_per", "stalled_wait_per", "stalled_math_pipe",
                        "stalled_mio_throttle", "stalled_not_selected", "local_load", "local_store", "lmem"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:100])
    for h, u in zip(hdr, units):
        if any(w in h for w in want) and "peak_sustained" not in h.split(".")[-1] and "per_second" not in h:
            print(f"   {h} = {d[h]} {u}")
