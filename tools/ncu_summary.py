"""Text summary of an `ncu --set full` report: one block per captured launch with the metrics the
roofline argument uses.  usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
print(f"# {rep}: {len(data)} launches captured with ncu --set full --clock-control none (cold-cache, serialised replays)")
for n, r in enumerate(data):
    print(f"\n[{n + 1}] {r[ix['Kernel Name']]}")
    rd = float(r[ix["dram__bytes_read.sum"]]) if "dram__bytes_read.sum" in ix else 0
    for w in want:
        if w in ix:
            print(f"    {w:72s} {r[ix[w]]:>16s} {units[ix[w]]}")
    try:
        def val(name):
            v, u = float(r[ix[name]]), units[ix[name]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        t = float(r[ix["gpu__time_duration.sum"]]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(units[ix["gpu__time_duration.sum"]], 1)
        tr = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        print(f"    {'traffic = dram read + write':72s} {tr / 1e9:16.3f} GB  -> {tr / t / 1e9:8.1f} GB/s under ncu")
    except Exception as e:
        print("    (traffic n/a)", e)
