"""Turn the ncu outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/:
    python tools/ncu_summaries.py r2 [note]
launch list   gpurun_out/launches_<round>*.csv        -> profiles/<round>_launches*_summary.csv  (per kernel: launches, total / average us, share)
full captures gpurun_out/prof_<round>_<kernel>.ncu-rep -> profiles/<round>_<kernel>_ncu_full.csv  (the metrics the roofline numbers are read from)"""
import collections, csv, glob, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
rnd = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""

METRICS = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.per_cycle_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= v:
            continue
        name = r[k].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[v].replace(",", "")) / 1000.0
    tot = sum(a[1] for a in agg.values())
    base = os.path.basename(path)[len("launches_"):-len(".csv")]
    out = os.path.join(ROOT, "profiles", f"{base}_launches_summary.csv")
    cmd_file = path[:-4] + ".cmd"
    cmd = open(cmd_file).read().strip() if os.path.exists(cmd_file) else "python bench.py ..."
    with open(out, "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {cmd} ({note}; cold-cache serialised times: compare shares)\n")
        fh.write("kernel,launches,total_us,avg_us,share_pct\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write(f"{n},{a[0]},{a[1]:.1f},{a[1] / a[0]:.2f},{100 * a[1] / tot:.1f}\n")
    print("wrote", out)


def full(path):
    name = os.path.basename(path)[len(f"prof_{rnd}_"):-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        print("empty", path)
        return
    hdr, units = rows[0], rows[1]
    out = os.path.join(ROOT, "profiles", f"{rnd}_{name}_ncu_full.csv")
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    cmd_file = path[:-8] + ".cmd"
    cmd = open(cmd_file).read().strip() if os.path.exists(cmd_file) else ""
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none --import-source on {cmd} ({note})\n")
        w = csv.writer(fh)
        w.writerow([f"{hdr[c]} [{units[c]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([r[c] for c in cols])
    print("wrote", out)


for p in sorted(glob.glob(os.path.join(GO, f"launches_{rnd}*.csv"))):
    launches(p)
for p in sorted(glob.glob(os.path.join(GO, f"prof_{rnd}_*.ncu-rep"))):
    full(p)
