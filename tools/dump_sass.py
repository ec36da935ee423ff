"""Regenerate profiles/<round>_sass_summary.csv (per-kernel instruction mix of libwhisper_b200.so) and the full SASS listings of the
hot kernels under profiles/sass/.  Needs only cuobjdump (no GPU):  python tools/dump_sass.py [round-prefix, default r1]"""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "whisper.coreml_b200", "libwhisper_b200.so")
PREFIX = sys.argv[1] if len(sys.argv) > 1 else "r1"
HOT = ("flash_attn_tc_kernel", "decoder_batch_kernel<1", "decoder_batch_kernel<5", "gemm_tcgen05", "layernorm_kernel<10>", "mel_frames_kernel",
       "sample_update_batch_kernel", "median_kernel", "dtw_kernel", "align_kernel", "attention_simt_kernel", "attention_qk_dump_kernel")
COLS = [("UTCHMMA(tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM/STTM(tcgen05.ld/st)", r"\b(LDTM|STTM)"), ("UTMALDG/UTMASTG(TMA tensor)", r"\bUTMA(LDG|STG)"),
        ("UBLKCP(cp.async.bulk)", r"\bUBLKCP"), ("HMMA(mma.sync)", r"\bHMMA"), ("LDGSTS(cp.async)", r"\bLDGSTS"), ("SYNCS(mbarrier)", r"\bSYNCS"),
        ("STL+LDL(local)", r"\b(STL|LDL)\b")]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
funcs, cur = [], None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = [m.group(1), []]
        funcs.append(cur)
    elif cur is not None:
        cur[1].append(line)

os.makedirs(os.path.join(ROOT, "profiles", "sass"), exist_ok=True)
for f in os.listdir(os.path.join(ROOT, "profiles", "sass")):
    if f.startswith(PREFIX + "_"):
        os.remove(os.path.join(ROOT, "profiles", "sass", f))
rows = []
for mangled, lines in funcs:
    name = demangle(mangled)
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"^b200::", "", short)
    short = short.replace("(int)", "").replace("(bool)", "")
    short = re.sub(r"\(.*$", "", short)
    insts = [l for l in lines if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]
    body = "\n".join(insts)
    rows.append([short, len(insts)] + [len(re.findall(p, body)) for _, p in COLS])
    if any(h in short for h in HOT):
        fn = PREFIX + "_" + re.sub(r"[^A-Za-z0-9]+", "_", short).strip("_") + ".sass"
        with open(os.path.join(ROOT, "profiles", "sass", fn), "w") as fh:
            fh.write("Function : " + mangled + "\n" + "\n".join(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in lines if l.strip()) + "\n")
with open(os.path.join(ROOT, "profiles", PREFIX + "_sass_summary.csv"), "w") as fh:
    fh.write("# cuobjdump -sass whisper.coreml_b200/libwhisper_b200.so : per-kernel instruction mix (sm_100a), regenerate with tools/dump_sass.py. "
             "Full listings of the hot kernels: profiles/sass/*.sass\n")
    fh.write("kernel,instructions," + ",".join(c for c, _ in COLS) + "\n")
    for r in rows:
        fh.write(",".join(str(x) for x in r) + "\n")
print(len(rows), "kernels;", len(os.listdir(os.path.join(ROOT, "profiles", "sass"))), "listings")
