"""Per-kernel summary of an `ncu --set full` report (run where ncu is installed, no GPU needed):
    python scripts_dev/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.csv
Launches of the same kernel instantiation are averaged; times in us, bytes in MB per launch."""
import collections, csv, io, re, subprocess, sys

METRICS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_subpipe_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_inst_pct", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("smem_KB", "launch__shared_mem_per_block_dynamic"),
    ("sm_clock_MHz", "sm__cycles_elapsed.avg.per_second"),
]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [(n, hdr.index(m)) for n, m in METRICS if m in hdr]
missing = [m for n, m in METRICS if m not in hdr]
agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("ub::", "")
    grid = r[hdr.index("Grid Size")]
    a = agg.setdefault(name, {"n": 0, "grid": grid, "v": collections.defaultdict(float)})
    a["n"] += 1
    for n, i in cols:
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            v = 0.0
        u = units[i]
        if n == "time_us" and u in ("ns", "nsecond"): v /= 1e3
        if n == "time_us" and u in ("ms", "msecond"): v *= 1e3
        if n.endswith("_MB"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        if n == "smem_KB":
            v *= {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(u, 1.0)
        if n == "sm_clock_MHz":
            v *= {"hz": 1e-6, "Khz": 1e-3, "Mhz": 1.0, "Ghz": 1e3}.get(u, 1.0)
        a["v"][n] += v
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "grid"] + [n for n, _ in cols])
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["v"]["time_us"]):
    w.writerow([name, a["n"], a["grid"]] + [f"{a['v'][n] / a['n']:.2f}" for n, _ in cols])
if missing:
    print("# metrics absent from this report: " + ", ".join(missing), file=sys.stderr)
