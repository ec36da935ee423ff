"""Turn the ncu outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/:
    python tools/ncu_summaries.py <tag> [note]      e.g.  r1f
launch list  -> profiles/<round>_launches_<tag>_summary.csv   (per kernel: launches, total / average us, share)
full captures -> profiles/<round>_<kernel>_<tag>_ncu_full.csv (the metrics the roofline numbers are read from)"""
import collections, csv, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
rnd = tag[:2]

METRICS = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.per_cycle_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum"]

def launches():
    path = os.path.join(GO, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
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
    out = os.path.join(ROOT, "profiles", f"{rnd}_launches_{tag}_summary.csv")
    with open(out, "w") as fh:
        cmd = os.environ.get("B200_LAUNCH_CMD", "python bench.py --steps 1 --warmup 1 --cpu-baseline 0 --long-clip 0 --sample-len 24")
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {cmd} ({note}; cold-cache serialised times: compare shares)\n")
        fh.write("kernel,launches,total_us,avg_us,share_pct\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write(f"{n},{a[0]},{a[1]:.1f},{a[1] / a[0]:.2f},{100 * a[1] / tot:.1f}\n")
    print("wrote", out)

def full(rep, name, cmd_note):
    path = os.path.join(GO, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = os.path.join(ROOT, "profiles", f"{rnd}_{name}_{tag}_ncu_full.csv")
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none --import-source on, {cmd_note} ({note})\n")
        w = csv.writer(fh)
        w.writerow([f"{hdr[c]} [{units[c]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([r[c] for c in cols])
    print("wrote", out)

launches()
full(f"prof_flash_{tag}.ncu-rep", "flash_attn", "-k regex:flash_attn -s 36 -c 1, bench.py --sample-len 2: one encoder layer's attention (1500 frames, 20 heads, 2 windows)")
full(f"prof_gemm_{tag}.ncu-rep", "gemm_tcgen05", "-k regex:gemm_tcgen05 -s 146 -c 4, bench.py --sample-len 2: the four GEMMs of one encoder layer at M = 3000 (qkv, out, mlp1+GELU, mlp2)")
full(f"prof_mega_{tag}.ncu-rep", "decoder_mega", "-k regex:decoder_mega -s 30 -c 1, bench.py --sample-len 24: one decoder1 token step of one lane (74 CTAs)")
