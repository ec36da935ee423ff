#!/usr/bin/env python
"""Headline benchmark: turbo (large-v3-turbo dims, random-init) beam-5 transcription RTFx on synthetic audio.

    python bench.py --gpus N --steps K --warmup W            # this framework on N B200s (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's PyTorch CPU path (oracle port) on host cores
    python bench.py --shard-file 60 --model large-v3 [--gpus N]   # BASELINE configs[4]: ONE long file, windows sharded r::N

One step = transcribe one 1-minute synthetic clip (2 fixed 30-s windows; BASELINE.json configs[3]) per GPU:
log-mel -> encoder -> crossKV -> decoder256 -> <=223 decoder1 steps with on-device beam search.  Multi-GPU is
weak scaling over independent clips (no collective in the loop).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is derived.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="turbo")
    ap.add_argument("--minutes", type=float, default=1.0, help="audio per step per GPU")
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--sample-len", type=int, default=224)
    ap.add_argument("--word-timestamps", action="store_true")
    ap.add_argument("--window-batch", type=int, default=8)
    ap.add_argument("--cpu-baseline", type=int, default=1, help="time the oracle port on host cores (N=1 only)")
    ap.add_argument("--long-clip", type=float, default=4.0, help="minutes of a second, longer clip reported as long_clip (0 = skip; N=1 only)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--shard-file", type=float, default=0.0,
                    help="minutes of ONE synthetic file whose fixed windows are sharded rank::world_size (BASELINE configs[4]; strong scaling)")
    ap.add_argument("--word-timestamps-pass", type=int, default=1, help="also time the workload with word_timestamps=True (N=1 only)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"          # /opt/skills/guides/B200_PROFILING.md


def weights_folder(name: str, seed: int):
    """Random-init weights of the named architecture (oracle recipe), exported once per box to /tmp."""
    from oracle import model as om
    from whisper_b200 import export
    dims = om.DIMS[name]
    folder = os.path.join(tempfile.gettempdir(), f"b200_bench_weights_{name}_{seed}")
    ckpt_path = os.path.join(folder, "ckpt.pt")
    done = os.path.join(folder, "DONE")
    if not os.path.exists(done):
        ckpt = om.init_weights(dims, seed)
        export.export_model(ckpt, dims, folder, fused=False)
        torch.save(ckpt, ckpt_path)
        open(done, "w").close()
    return dims, folder, ckpt_path


# ---- algorithmic work ---------------------------------------------------------
def encoder_flops(dims) -> float:
    d, le, m = dims.n_audio_state, dims.n_audio_layer, dims.n_mels
    return 2 * 3000 * 3 * m * d + 2 * 1500 * 3 * d * d + le * (24 * 1500 * d * d + 4 * 1500 * 1500 * d)


def decoder1_bytes(dims, bs: int, t: float, windows: int = 1) -> float:
    """SURVEY 8d for ONE launch of the step kernel that advances `windows` windows: the weights once, everything else per window"""
    d, ld, v = dims.n_text_state, dims.n_text_layer, dims.n_vocab
    weights = 2 * (ld * (14 * d * d + 16 * d) + 2 * d + v * d)
    per_window = 4 * ld * 1500 * d + 4 * ld * bs * (t + 1) * d + 4 * ld * bs * d + 4 * bs * d + 4 * bs * v
    return weights + windows * per_window


def lane_plan(n_windows: int, nb: int):
    """(lanes, windows per lane) of b200DecodeWindows for one batch of windows (csrc/api_batch.cu: decode_windows_batch)"""
    lanes = min(int(os.environ.get("B200_DECODE_LANES", "8")), 8, n_windows)
    per_lane = min(int(os.environ.get("B200_BATCH_WINDOWS", "8")), max(1, 40 // nb), -(-n_windows // lanes))
    return lanes, per_lane


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from this round's committed ncu capture
    (profiles/r2_<kernel>_ncu_full.csv, written by tools/ncu_summaries.py), or None"""
    import csv
    path = os.path.join(ROOT, "profiles", f"r2_{kernel}_ncu_full.csv")
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r and not r[0].startswith("#")]
    if len(rows) < 2:
        return None
    hdr = rows[0]
    total, n = 0.0, 0
    for r in rows[1:]:
        t = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            col = next((i for i, h in enumerate(hdr) if h.startswith(name + " [")), None)
            if col is None:
                return None
            t += float(r[col].replace(",", "")) * scale.get(hdr[col].split("[")[1].rstrip("]"), 1.0)
        total += t
        n += 1
    return total / n if n else None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        # NVML in-process (a sample takes ~0.1 ms) when nvidia-ml-py is importable, else one nvidia-smi process per sample:
        # the timed region of the default run is < 1 s, and a cold nvidia-smi start alone can take longer than that
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")]
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx), ""] + ["Active" if r & b else "Not Active" for b, _ in bits])
                time.sleep(0.05)
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference(dims, ckpt_path, n_mels, sample_len, beam, cap_steps, audio_seconds, n_windows):
    """The reference's PyTorch CPU path (oracle port of whisper/encoder.py, decoder.py, decoding.py, fp32) on the host
    cores: one 30-s window, decoder loop capped at `cap_steps`, extrapolated linearly to `sample_len` steps and to the
    clip's `n_windows` windows."""
    from oracle import audio as oa, decoding as od, model as om, synth
    # all the host cores the process may use (torch's default without OMP_NUM_THREADS): torchrun exports OMP_NUM_THREADS=1,
    # which would leave the CPU path single-threaded
    try:
        logical = len(os.sched_getaffinity(0))
    except Exception:
        logical = os.cpu_count() or 1
    torch.set_num_threads(max(torch.get_num_threads(), logical, 1))
    ckpt = torch.load(ckpt_path)
    orc = om.OracleModel(dims, ckpt)
    audio = synth.noise_audio(1, 480000)
    t0 = time.perf_counter()
    mel = oa.log_mel_spectrogram(audio, n_mels, padding=480000)[:, :3000].contiguous()
    t_mel = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.encode(mel)
    t_enc = time.perf_counter() - t0
    sp = od.Specials.load(dims.n_vocab)
    t0 = time.perf_counter()
    r1 = od.decode_window(orc, None, sp, od.Options(sample_len=1, beam_size=beam))
    t_prefill = time.perf_counter() - t0
    t0 = time.perf_counter()
    rc = od.decode_window(orc, None, sp, od.Options(sample_len=cap_steps, beam_size=beam))
    t_cap = time.perf_counter() - t0
    per_step = max(t_cap - t_prefill, 0.0) / max(rc.steps - 1, 1)
    finished_early = rc.steps < cap_steps
    full_steps = rc.steps if finished_early else sample_len
    t_window = t_mel + t_enc + t_prefill + per_step * (full_steps - 1)
    return {"rtfx": audio_seconds / (t_window * n_windows), "t_window_s": t_window, "t_mel_s": t_mel, "t_encoder_s": t_enc,
            "t_prefill_s": t_prefill, "t_step_s": per_step, "steps_timed": rc.steps, "steps_extrapolated": full_steps,
            "cores": torch.get_num_threads()}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    from oracle import model as om, synth
    dims = om.DIMS[a.model]
    n_samples = int(a.minutes * 60 * 16000)
    audio_seconds = n_samples / 16000.0
    n_windows = (n_samples + 480000 - 1) // 480000
    workload = (f"{a.model} dims random-init (seed {a.seed}), beam_size={a.beam}, {a.minutes:g}-min synthetic audio "
                f"(randn*0.1, seed 1+rank) = {n_windows} fixed 30-s windows per GPU per step, sample_len={a.sample_len}, "
                f"condition_on_previous_text=False, temperature 0")
    config = {"workload": workload, "windows_per_step_per_gpu": n_windows, "beam_size": a.beam, "sample_len": a.sample_len,
              "word_timestamps": bool(a.word_timestamps), "l2": "working set (>2 GB of weights per step) exceeds the 126 MB L2",
              "decode_lanes": "windows are spread over up to 8 concurrent decode lanes (B200_DECODE_LANES); a lane advances its windows in "
                              "one batched step kernel (one window per lane up to 8 windows per batch)"}

    if a.impl == "reference":
        if rank != 0:
            return
        _, _, ckpt_path = weights_folder(a.model, a.seed)
        vals, walls = [], []
        cap = min(25, a.sample_len)                                      # bounded sample: 24 decoder1 steps of one window per step
        for i in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            r = cpu_reference(dims, ckpt_path, dims.n_mels, a.sample_len, a.beam, cap, audio_seconds, n_windows)
            if i >= a.warmup:
                vals.append(r)
                walls.append(time.perf_counter() - t0)
        v = statistics.mean(x["rtfx"] for x in vals)
        last = vals[-1]
        sample = (f"per step 1 of {n_windows} windows: mel + encoder + crossKV + decoder256 x{a.beam} + {last['steps_timed']-1} decoder1 "
                  f"steps + beam search, fp32 torch CPU (oracle port of the reference's use_coreml=False path); RTFx extrapolates the "
                  f"measured per-step decoder time to {last['steps_extrapolated']} steps and the window to {n_windows} windows; "
                  f"ms_per_step is the wall time of the sample actually run")
        print(json.dumps({"impl": "reference", "metric": "rtfx", "value": v, "unit": "audio_s/wall_s", "n_gpus": a.gpus,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * statistics.mean(walls), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "extrapolated": True,
                          "cpu_baseline": {"value": v, "unit": "audio_s/wall_s", "cores": last["cores"], "kind": "port", "sample": sample,
                                           "detail": {k: last[k] for k in ("t_mel_s", "t_encoder_s", "t_prefill_s", "t_step_s")}},
                          "e2e": {"value": v, "unit": "audio_s/wall_s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------------------------------------------------------- B200 arm
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if rank == 0:
        _, folder, ckpt_path = weights_folder(a.model, a.seed)
    if world > 1:
        dist.barrier()
    _, folder, ckpt_path = weights_folder(a.model, a.seed)
    from whisper_b200 import _lib
    from whisper_b200.model import ModelDimensions, WhisperB200
    from whisper_b200.transcribe import transcribe
    model = WhisperB200(ModelDimensions(**dims.as_dict()), folder, device=local, beam_slots=max(a.beam, 1)).load()
    lib = model.lib
    audio_host = synth.noise_audio(1 + rank, n_samples).pin_memory()
    audio_dev = audio_host.cuda()
    kw = dict(beam_size=a.beam or None, sample_len=a.sample_len, word_timestamps=a.word_timestamps, window_batch=a.window_batch)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if a.shard_file > 0:
        # BASELINE configs[4]: ONE long file, its fixed 30-s windows sharded rank::world_size (whisper/transcribe.py:276-290 loops the
        # windows of a file); strong scaling: the file is the same at every N.  value = file seconds / max-over-ranks device time.
        from whisper_b200.transcribe import gather_sharded
        n_file = int(a.shard_file * 60 * 16000)
        file_host = synth.noise_audio(1, n_file).pin_memory()            # every rank holds the whole file (the log-mel needs its global max)
        file_dev = file_host.cuda()
        skw = dict(kw, rank=rank, world_size=world)
        for _ in range(max(a.warmup, 1)):
            transcribe(model, file_dev, **skw)
        vals = {}
        for name, src in (("dev", file_dev), ("e2e", file_host)):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            model.stage_times_ms(reset=True)
            e0.record()
            for _ in range(a.steps):
                res = transcribe(model, src, **skw)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            vals[name] = float(t[0])
        merged = gather_sharded(res, world)
        stages = model.stage_times_ms()
        if rank == 0:
            total_windows = (n_file + 480000 - 1) // 480000
            assert merged["windows"] == total_windows, (merged["windows"], total_windows)
            secs = n_file / 16000.0
            print(json.dumps({"metric": "rtfx", "value": secs * a.steps / (vals["dev"] / 1000.0), "unit": "audio_s/wall_s", "n_gpus": world,
                              "steps": a.steps, "warmup": a.warmup, "ms_per_step": vals["dev"] / a.steps, "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                              "config": {"workload": f"{a.model} dims random-init, beam_size={a.beam}, ONE {a.shard_file:g}-min synthetic file = "
                                                     f"{total_windows} fixed 30-s windows sharded rank::world_size over {world} GPU(s) "
                                                     f"(BASELINE configs[4]), sample_len={a.sample_len}", "windows": total_windows,
                                         "windows_rank0": res["windows"], "window_batch": a.window_batch},
                              "e2e": {"value": secs * a.steps / (vals["e2e"] / 1000.0), "unit": "audio_s/wall_s",
                                      "h2d_bytes_per_step": 4 * n_file, "d2h_bytes_per_step": res["windows"] * (max(a.beam, 1) * 449 * 4 + 44)},
                              "stage_ms_per_step_rank0": {k: v / a.steps for k, v in stages.items()},
                              "segments": len(merged["segments"]), "tokens": sum(len(s["tokens"]) for s in merged["segments"])}))
        if world > 1:
            dist.destroy_process_group()
        return

    def timed(audio, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.b200KernelLaunchCount()
        model.stage_times_ms(reset=True)
        t0 = time.perf_counter()
        e0.record()
        res = None
        for _ in range(steps):
            res = transcribe(model, audio, **kw)          # synchronous: returns with the tokens in host memory
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1000.0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), res, lib.b200KernelLaunchCount() - l0, model.stage_times_ms()

    for _ in range(max(a.warmup, 0)):
        transcribe(model, audio_dev, **kw)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, wall_dev, res, launches, stages = timed(audio_dev, a.steps)
    ms_e2e, wall_e2e, res2, _, _ = timed(audio_host, a.steps)
    sampler.stop_flag = True
    value = world * audio_seconds * a.steps / (ms_dev / 1000.0)
    e2e = world * audio_seconds * a.steps / (ms_e2e / 1000.0)
    if rank != 0:
        return
    hbm, tf, which = peaks()
    n_tok = sum(len(s["tokens"]) for s in res["segments"])
    out = {"metric": "rtfx", "value": value, "unit": "audio_s/wall_s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic", "config": config,
           "e2e": {"value": e2e, "unit": "audio_s/wall_s", "h2d_bytes_per_step": 4 * n_samples,
                   "d2h_bytes_per_step": n_windows * (max(a.beam, 1) * 449 * 4 + max(a.beam, 1) * 8 + 4)},
           "gpu_launches": int(launches), "clocks": sampler.summary(), "tokens_per_step": n_tok,
           "stage_ms_per_step": {k: v / a.steps for k, v in stages.items()}}
    # decoder1 roofline: one launch of decoder_batch_kernel advances the windows of one lane by one token.  achieved = algorithmic
    # bytes of a launch (the weights once + per window cross K/V, self K/V, I/O) x launches / CUDA-event time of the
    # whole decoder1 loop (sampling kernels and launch gaps included; the lanes run concurrently, so this is their aggregate rate)
    dec_steps = sum(max(x - 1, 0) for x in res["decode_steps"])
    if dec_steps:
        nb = max(a.beam, 1)
        lanes, per_lane = lane_plan(min(n_windows, a.window_batch), nb)
        per_step_ms = stages["decoder1"] / a.steps / dec_steps           # per window-step
        n_init = len(model.specials.sot_sequence)
        t_mean = sum(n_init + (x - 1) / 2.0 for x in res["decode_steps"]) / len(res["decode_steps"])
        by_launch = decoder1_bytes(dims, nb, t_mean, per_lane)
        ach = by_launch / per_lane / (per_step_ms * 1e-3) / 1e9
        out["roofline"] = {"kernel": "decoder_batch_kernel: one persistent launch per decoder1 token step of a lane (LN + 7 GEMVs per layer, "
                                     "self / cross attention, vocabulary projection for every window x beam row of the lane) + the "
                                     "sampling kernel that follows it",
                           "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                           "traffic": ncu_traffic("decoder_batch"),
                           "peak_source": which, "bytes_per_launch": by_launch, "windows_per_launch": per_lane, "lanes": lanes,
                           "us_per_window_step": per_step_ms * 1e3, "us_per_step": per_step_ms * 1e3, "steps": dec_steps,
                           "note": f"algorithmic bytes of one launch (weights once + {per_lane} window(s) of cross K/V, self K/V and I/O at the "
                                   f"mean text_offset) per window, divided by the CUDA-event time of the decoder1 loop per window-step; "
                                   f"{lanes} lane(s) run concurrently on {lanes} disjoint SM groups, so this is their aggregate HBM rate; "
                                   f"traffic = dram bytes of ONE launch from profiles/r2_decoder_batch_ncu_full.csv (null if absent)"}
    enc_ms = stages["encoder"] / a.steps / n_windows
    if enc_ms > 0:
        fl = encoder_flops(dims)
        out["encoder_roofline"] = {"bound": "tensor", "achieved": fl / (enc_ms * 1e-3) / 1e12, "peak": tf, "unit": "TFLOP/s",
                                   "frac": fl / (enc_ms * 1e-3) / 1e12 / tf, "flops_per_window": fl, "ms_per_window": enc_ms,
                                   "peak_source": which + " (sustained cuBLAS bf16)"}
    if world == 1 and a.long_clip > 0:
        # the same path on a longer clip: more independent windows -> more concurrent decode lanes (not the headline workload)
        nl = int(a.long_clip * 60 * 16000)
        long_audio = synth.noise_audio(101, nl).cuda()
        for _ in range(2):                                   # (the second pass captures the encoder graph of this window count)
            transcribe(model, long_audio, **kw)
        ms_l, _, res_l, _, st_l = timed(long_audio, 2)
        wl = (nl + 480000 - 1) // 480000
        ds = sum(max(x - 1, 0) for x in res_l["decode_steps"])
        t_mean_l = sum(len(model.specials.sot_sequence) + (x - 1) / 2.0 for x in res_l["decode_steps"]) / len(res_l["decode_steps"])
        by_l = decoder1_bytes(dims, max(a.beam, 1), t_mean_l, 1)
        ach_l = by_l / (st_l["decoder1"] / 2 / ds * 1e-3) / 1e9 if ds else 0.0
        out["long_clip"] = {"minutes": a.long_clip, "windows_per_step": wl, "decode_lanes": min(8, wl), "value": (nl / 16000.0) * 2 / (ms_l / 1000.0),
                            "unit": "audio_s/wall_s", "ms_per_step": ms_l / 2,
                            "decoder1_hbm_gbs": ach_l, "decoder1_hbm_frac": ach_l / hbm,
                            "encoder_tflops": encoder_flops(dims) / (st_l["encoder"] / 2 / wl * 1e-3) / 1e12,
                            "note": "same transcribe() call on a longer clip (window_batch 8): algorithmic decoder1 bytes per window-step / "
                                    "CUDA-event time of the decoder1 loop per window-step, aggregate over the concurrent lanes"}
    if world == 1 and a.word_timestamps_pass and not a.word_timestamps:
        # the same workload with word_timestamps=True (cross-attention of the alignment heads, median filter, DTW per window)
        kw_w = dict(kw, word_timestamps=True)
        transcribe(model, audio_dev, **kw_w)
        kw_saved = dict(kw)
        kw.update(kw_w)
        ms_w, _, res_w, _, st_w = timed(audio_dev, max(2, a.steps // 2))
        kw.clear(); kw.update(kw_saved)
        n_w = max(2, a.steps // 2)
        out["word_timestamps"] = {"value": audio_seconds * n_w / (ms_w / 1000.0), "unit": "audio_s/wall_s", "ms_per_step": ms_w / n_w,
                                  "stage_ms_per_step": {k: v / n_w for k, v in st_w.items()},
                                  "words": sum(len(s.get("words", [])) for s in res_w["segments"]),
                                  "relative_to_value": (audio_seconds * n_w / (ms_w / 1000.0)) / value}
    if a.cpu_baseline and world == 1:
        r = cpu_reference(dims, ckpt_path, dims.n_mels, a.sample_len, a.beam, a.sample_len, audio_seconds, n_windows)   # ~10 s: one full window
        out["cpu_baseline"] = {"value": r["rtfx"], "unit": "audio_s/wall_s", "cores": r["cores"], "kind": "port",
                               "sample": f"1 of {n_windows} windows: mel + encoder + crossKV + decoder256 x{a.beam} + {r['steps_timed']-1} "
                                         f"decoder1 steps (fp32 torch CPU oracle), extrapolated to {r['steps_extrapolated']} steps x {n_windows} windows",
                               "detail": {k: r[k] for k in ("t_mel_s", "t_encoder_s", "t_prefill_s", "t_step_s")}}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
