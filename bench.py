"""bench.py -- clips/s of the CQT + PitchClassNet forward hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle port on the host cores

A step = one pass of the hot path (CQT -> PitchClassNet forward -> argmax decode) over one batch
of synthetic clips per GPU.  Workload (BASELINE.json configs[1]): 256 mono fp32 clips per GPU,
48 kHz x 30 s (1,440,000 samples, hop 9600 -> 151 frames, 288 bins), PitchClassNet at the
train_model.py defaults (+ genre head), seeded weights with randomised BatchNorm statistics
(tests/golden/weights_seed0.npz).  `value`: inputs resident in HBM; `e2e`: pinned HOST audio in,
host predictions out through ake_estimate_host_f32 (H2D + D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, SECONDS_STD, FRAMES, OCTAVES = 48000, 30, 5, 8
METRIC = "clips/s CQT+PitchClassNet fwd"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--seconds", type=float, default=SECONDS_STD, help="clip duration")
    ap.add_argument("--no-genre", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather-every", type=int, default=4,
                    help="multi-GPU: all-gather the (clips, 35) result rows once per this many steps (rows of the steps in between wait "
                         "in a device buffer); 1 = after every step")
    ap.add_argument("--no-graph", action="store_true", help="--train: launch the step's kernels one by one instead of as a CUDA graph")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE configs[3]: large-batch sweep (1k/2k/4k/8k/16k long clips of 240 s, global batch sharded over "
                         "--gpus ranks, processed in resident sub-batches, result rows all-gathered); one JSON line per point")
    ap.add_argument("--sweep-batches", default="1024,2048,4096,8192,16384")
    ap.add_argument("--sweep-seconds", type=float, default=240.0)
    ap.add_argument("--sweep-sub", type=int, default=64, help="--sweep: clips resident per sub-batch (64 x 46 MB = 2.9 GB)")
    ap.add_argument("--train", action="store_true",
                    help="BASELINE configs[4] instead of the headline metric: training step (fwd + loss + bwd, batch 8 per GPU, "
                         "train-mode BN) with ONE flat gradient all-reduce over NCCL; prints its own JSON line")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def golden_weights(genre: bool):
    import numpy as np
    import torch
    w = np.load(os.path.join(ROOT, "tests", "golden", "weights_seed0.npz"))
    return {k: torch.from_numpy(w[k]) for k in w.files if genre or not k.startswith("genre_classifier.")}


# ------------------------------------------------------------------------------------ clock sampling
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU oracle arm
def cpu_oracle_step(clips, sd64, genre, pool):
    """One reference-side step on the host: oracle CQT per clip (thread pool, as KeyDataset.py:127 preloads)
    then the float64 forward (the reference's dtype, train_model.py:105) + argmax decode."""
    import numpy as np
    import torch
    from oracle import cqt_port, pcn_port
    mels = list(pool.map(lambda c: cqt_port.cqt_logmag(c, SR, FRAMES, OCTAVES, dtype=np.float32), clips))
    T = max(m.shape[-1] for m in mels)
    x = torch.from_numpy(np.stack([np.pad(m, ((0, 0), (0, 0), (0, T - m.shape[-1]))) for m in mels]))
    with torch.no_grad():
        out = pcn_port.pcn_forward(sd64, x, torch.tensor([m.shape[-1] for m in mels]))
        ids = pcn_port.decode(*out)
    return out, ids


def time_cpu_oracle(n_clips, n_samples, steps, warmup, genre):
    import concurrent.futures as cf
    import torch
    from audio_key_estimation_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd64 = {k: v.double() for k, v in golden_weights(genre).items() if v.is_floating_point()}
    clips = [synth.synth_clip(i, n_samples, SR).numpy() for i in range(n_clips)]
    with cf.ThreadPoolExecutor(max_workers=cores) as pool:
        for _ in range(warmup):
            cpu_oracle_step(clips, sd64, genre, pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_oracle_step(clips, sd64, genre, pool)
        dt = time.perf_counter() - t0
    return n_clips * steps / dt, dt / steps, cores


# ONE protocol for both CPU legs (the in-arm `cpu_baseline` and `--impl reference`): the same function on the same bounded
# sample -- CPU_CLIPS clips per step (enough to occupy the thread pool), one warm-up step, then timed steps.
CPU_CLIPS = 8


def cpu_sample_text(n_clips, seconds, cores):
    return (f"{n_clips} clips/step of the b200 arm's workload ({seconds:g} s @ {SR} Hz): oracle CQT (numpy fp32, thread pool of {cores}) "
            f"+ float64 forward (torch CPU, {cores} threads) + argmax decode")


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference is pure Python
    (models.py + librosa) and cannot travel to the GPU box, so this times the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_samples = int(round(args.seconds * SR))
    genre = not args.no_genre
    # bounded sample: CPU_CLIPS clips per step, shrunk only if the whole --steps/--warmup run would pass ~150 s
    t0 = time.perf_counter()
    time_cpu_oracle(1, n_samples, 1, 0, genre)
    t_clip = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_clips = int(max(1, min(args.batch, CPU_CLIPS, budget / max(t_clip / min(cores, 8), 1e-3))))
    value, step_s, cores = time_cpu_oracle(n_clips, n_samples, args.steps, args.warmup, genre)
    sample = cpu_sample_text(n_clips, args.seconds, cores)
    cfg = workload_config(args, genre)
    # what this arm really ran per step (the b200 arm's per-GPU batch is echoed for the driver's config match only)
    cfg.update({"reference_clips_per_step": n_clips, "reference_ranks": 1,
                "reference_note": "CPU arm: rank 0 only, all host cores, a bounded sample of the workload per step; it does not scale with --gpus "
                                  "(ratios against it are meaningful at N = 1)"})
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, genre):
    n_samples = int(round(args.seconds * SR))
    return {"workload": f"configs[1]: {args.batch} synthetic mono clips per GPU, {args.seconds:g} s @ {SR} Hz "
                        f"({n_samples} samples, hop {SR // FRAMES}, {1 + n_samples // (SR // FRAMES)} frames x {36 * OCTAVES} bins), "
                        f"CQT + PitchClassNet(288,12,2,7) train_model.py defaults{' + genre head' if genre else ''} + decode, eval-mode BN",
            "clips_per_gpu": args.batch, "global_batch": args.batch * args.gpus, "clip_seconds": args.seconds, "sr": SR,
            "l2_policy": "inputs larger than L2 (audio batch >> 126 MB), no flush", "parallelism": f"dp{args.gpus} (batch sharded, logit all-gather" + (f" every {max(1, args.gather_every)} steps)" if args.gpus > 1 else ")")}


# ------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import audio_key_estimation_b200 as ake
    from audio_key_estimation_b200 import _lib, distributed as akd, synth, workload

    rank, local, world = akd.init_from_env("nccl")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (python -m torch.distributed.run --nproc-per-node N bench.py ...)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: keep each rank's pinned host buffers in the memory next to its GPU (multi-rank runs only, so the
    # N = 1 CPU baseline keeps every host core)
    if world > 1 and not os.environ.get("AKE_NO_NUMA_BIND"):
        akd.bind_to_gpu_numa_node(local)
    genre = not args.no_genre
    peaks = load_peaks()
    B, n_samples = args.batch, int(round(args.seconds * SR))

    net = ake.PitchClassNet(36 * OCTAVES, 12, 2, 7, opt=ake.default_opt(genre=genre))
    sd = golden_weights(genre)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    est = ake.KeyEstimator(net, SR, FRAMES)
    T = est.frames(n_samples)

    lo, _ = akd.shard_range(B * world, rank, world)
    audio = torch.empty((B, n_samples), dtype=torch.float32, device=dev)
    synth.synth_batch(lo, B, n_samples, SR, device=dev, out=audio)
    torch.cuda.synchronize()

    # Result rows of up to `gather_every` steps wait in a device buffer and cross NVLink in ONE all-gather (the collective is
    # latency-bound: 36 KB per rank and step).  Measured on 8 GPUs: a gather after every step costs ~60 us
    # per 2.4 ms step on the compute stream; issued asynchronously on NCCL's own stream it costs MORE (~120 us: its CTAs take
    # SMs from the one-CTA-per-SM persistent kernels while they wait for the slowest rank); one gather per 4 steps ~15 us per step.
    K_g = max(1, args.gather_every) if world > 1 else 1
    acc = torch.empty((K_g, B, akd.ROW), dtype=torch.float32, device=dev) if world > 1 else None
    n_acc = [0]
    last_table = [None]

    def flush():
        if world > 1 and n_acc[0]:
            k = n_acc[0]
            # (k * B) rows per rank -> (world, k, B, 35); the table of the latest step is [:, k - 1]
            got = akd.gather_rows(acc[:k].reshape(k * B, akd.ROW), k * B * world).reshape(world, k, B, akd.ROW)
            last_table[0] = got[:, k - 1].reshape(B * world, akd.ROW)
            n_acc[0] = 0

    def device_step():
        # CQT, then forward + decode in one call that writes the (clips, 35) result rows
        rows, ids = est.estimate_device_rows(audio)
        if world == 1:
            last_table[0] = rows
        else:
            acc[n_acc[0]].copy_(rows)
            n_acc[0] += 1
            if n_acc[0] == K_g:
                flush()
        return last_table[0], ids

    def drain():
        flush()
        return last_table[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput ("value")
    for _ in range(max(args.warmup, 3)):
        device_step()
    drain()
    barrier()
    lib = _lib.lib()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.ake_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        _, ids = device_step()
    table = drain()  # the last step's gather is inside the timed region
    ev1.record()
    barrier()
    if table.shape != (B * world, akd.ROW):
        raise SystemExit("gathered result table has the wrong shape")
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = int(lib.ake_launch_count(1))
    value = B * world * args.steps / (ms_total * 1e-3)
    # ---- the same K steps once more with the library's per-section CUDA events on (roofline / stage breakdown); kept out of
    # the timed region above: ~30 event records per step between the kernels are not part of the product path
    time.sleep(0.25)  # let the board's power budget recover: a back-to-back second pass runs ~7 % slower under sw_power_cap
    device_step()
    _lib.profile_enable(True)
    barrier()
    for _ in range(args.steps):
        device_step()
    drain()
    barrier()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)

    # ---- end to end through the C ABI with host buffers ("e2e")
    e2e, e2e_i16 = None, None
    if not args.no_e2e:
        host_audio = torch.empty((B, n_samples), dtype=torch.float32).pin_memory()
        host_audio.copy_(audio)
        out_bufs = None
        for _ in range(2):
            out_bufs = est.estimate_host(host_audio, out=out_bufs)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            out_bufs = est.estimate_host(host_audio, out=out_bufs)
            if world > 1:
                rows = akd.pack_rows(out_bufs["key"], out_bufs["tonic"], out_bufs["genre"]).to(dev)
                akd.gather_rows(rows, B * world)
        ev1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3  # estimate_host ends with a stream sync: wall ~ device time
        ms_e2e = max_over_ranks(max(ev0.elapsed_time(ev1), wall_ms))
        d2h = sum(v.numel() * v.element_size() for v in out_bufs.values() if v is not None)
        e2e = {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "clips/s",
               "h2d_bytes_per_step": int(host_audio.numel() * 4), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e2e / args.steps, "api": "KeyEstimator.estimate_host -> ake_estimate_host_f32 (pinned host audio)"}
        # parity guard: host-buffer path == device-resident path
        for j in range(3 if genre else 2):
            if not torch.equal(out_bufs["ids"][j].to(dev), ids[j]):
                raise SystemExit("e2e ids differ from the device-resident path")
        # ---- the same call fed with 16-bit PCM (what the reference's .wav files hold; normalised on the device exactly as
        # torchaudio.load does, KeyDataset.py:478-481): half the PCIe bytes.  Reported beside, not instead of, the fp32 `e2e`.
        host_pcm = torch.empty((B, n_samples), dtype=torch.int16).pin_memory()
        host_pcm.copy_((audio.clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16))
        out16 = None
        for _ in range(2):
            out16 = est.estimate_host(host_pcm, out=out16)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            out16 = est.estimate_host(host_pcm, out=out16)
            if world > 1:
                rows = akd.pack_rows(out16["key"], out16["tonic"], out16["genre"]).to(dev)
                akd.gather_rows(rows, B * world)
        ev1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms16 = max_over_ranks(max(ev0.elapsed_time(ev1), wall_ms))
        e2e_i16 = {"value": B * world * args.steps / (ms16 * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": int(host_pcm.numel() * 2),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": ms16 / args.steps,
                   "api": "KeyEstimator.estimate_host -> ake_estimate_host_i16 (pinned 16-bit PCM host audio)"}

    clocks = sampler.stop() if rank == 0 else None  # sampled across both timed regions (device-resident and end to end)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline bookkeeping (algorithmic work per step per GPU)
    macs_clip = workload.pcn_macs({k: tuple(v.shape) for k, v in sd.items()}, 36 * OCTAVES, T)
    p2p_macs_clip = workload.p2p_macs(36 * OCTAVES, T)
    cqt_bytes_clip = workload.cqt_algorithmic_bytes(n_samples, 36 * OCTAVES, T)
    steps = args.steps

    def sec(tag):
        ms, n = prof.get(tag, (0.0, 0))
        return ms / steps, n // steps if steps else 0

    p2p_ms, p2p_n = sec("pcn.p2p")
    pcn_ms, _ = sec("pcn.total")
    cqt_ms, _ = sec("cqt.total")
    dec_ms, dec_n = sec("cqt.decimate")
    bank_ms, bank_n = sec("cqt.bank")
    p2p_tflops = 2.0 * p2p_macs_clip * B / (p2p_ms * 1e-3) / 1e12 if p2p_ms else None
    # DRAM bytes per launch of the dominant kernel come from the committed ncu capture (profiles/), valid for the default
    # workload only; never measured live (a number taken under a profiler is not a bench value)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
    if os.path.exists(tpath) and B == 256 and abs(args.seconds - SECONDS_STD) < 1e-9 and genre:
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    # The scheme's own ceiling: operands are fp16 hi/lo pairs (22-bit) and the layer has 8 channels, so a block of 122 anchors is
    # 11 tcgen05.mma of N = 56 (7 time-tap phases x 8 channels), K = 16: seven (x_hi | x_lo) . W_hi row taps and four x_hi . W_lo
    # row-tap pairs, each bound by its 4 KB operand read out of shared memory (32 + N/4 = 46 SM cycles; tools/umma_rate.cu).
    sm_hz = ((clocks or {}).get("sm_max_mhz") or 1965.0) * 1e6
    p2p_blocks = B * 3 * 288 * T / 122.0   # output positions of the 3 convs / anchors per MMA block
    p2p_floor_ms = p2p_blocks * 11 * 46 / (148 * sm_hz) * 1e3
    scheme_ceiling = (2.0 * p2p_macs_clip * B / (p2p_floor_ms * 1e-3) / 1e12) / peaks["bf16_tflops_sustained"]
    roofline = {
        "kernel": "pcn.p2p: 3 x Conv2d 7x7 circular (pitch,time) + BN + LeakyReLU (64.6% of the reference forward's MACs); "
                  "4 launches: first conv split into its 36-periodic part + mel part (pcn_p2p1.cuh), then 2 x p2p_umma_kernel",
        "bound": "tensor", "achieved": p2p_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": (p2p_tflops / peaks["bf16_tflops_sustained"]) if p2p_tflops else None, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_flops_per_launch": 2.0 * p2p_macs_clip * B / max(1, p2p_n),
        "peak_source": f"{peaks['source']} bf16 dense, sustained (kernel timed inside a long step)",
        "launches_per_step": p2p_n, "ms_per_step": p2p_ms,
        "algorithmic_flops_per_step": 2.0 * p2p_macs_clip * B,
        "scheme_ceiling": scheme_ceiling,
        "scheme_ceiling_note": "operand-read floor of the fp16 hi/lo shift-GEMM at 8 channels (11 x N=56,K=16 MMAs of 46 cycles per 122 anchors) "
                               "expressed as a fraction of the dense bf16 peak: frac / scheme_ceiling = share of the achievable",
        "frac_of_scheme_ceiling": ((p2p_tflops / peaks["bf16_tflops_sustained"]) / scheme_ceiling) if p2p_tflops else None,
        "dtype_note": "fp16 hi/lo 3-product (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo), 22-bit operands, fp32 accumulate in TMEM",
    }
    stages = {
        "cqt": {"ms_per_step": cqt_ms, "bound": "hbm", "algorithmic_bytes_per_clip": cqt_bytes_clip,
                "achieved_gbs": (cqt_bytes_clip * B / (cqt_ms * 1e-3) / 1e9) if cqt_ms else None, "peak_gbs": peaks["hbm_gbs"],
                "frac": (cqt_bytes_clip * B / (cqt_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if cqt_ms else None,
                "decimate_ms": dec_ms, "bank_ms": bank_ms},
        "pcn": {"ms_per_step": pcn_ms, "bound": "tensor", "algorithmic_macs_per_clip": macs_clip,
                "achieved_tflops": (2.0 * macs_clip * B / (pcn_ms * 1e-3) / 1e12) if pcn_ms else None,
                "peak_tflops": peaks["bf16_tflops_sustained"],
                "frac": (2.0 * macs_clip * B / (pcn_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]) if pcn_ms else None,
                "sections_ms": {k: v[0] / steps for k, v in prof.items() if k.startswith("pcn.") and k != "pcn.total"}},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cv, cstep, cores = time_cpu_oracle(CPU_CLIPS, n_samples, 12, 1, genre)  # same protocol as --impl reference; ~10 s
        cpu_baseline = {"value": cv, "unit": "clips/s", "cores": cores, "kind": "port",
                        "sample": cpu_sample_text(CPU_CLIPS, args.seconds, cores) + "; 1 warm-up + 12 timed steps"}

    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, genre),
        "clocks": clocks, "e2e": e2e, "e2e_i16": e2e_i16, "gpu_launches": launches, "roofline": roofline, "stages": stages,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_sweep(args):
    """BASELINE configs[3]: global batches of 1k..16k LONG clips (240 s = 11,520,000 samples, 1201 frames) sharded over the
    ranks in contiguous clip ranges.  A rank keeps ONE sub-batch of --sweep-sub synthetic clips resident in HBM (64 x 46 MB)
    and runs it once per sub-batch of its shard -- 2048 clips x 46 MB never sit in HBM at once -- writing each sub-batch's
    result rows into its shard's (clips, 35) table, which is all-gathered once per step.  One JSON line per batch size."""
    import torch
    import torch.distributed as dist

    import audio_key_estimation_b200 as ake
    from audio_key_estimation_b200 import _lib, distributed as akd, synth

    rank, local, world = akd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    genre = not args.no_genre
    net = ake.PitchClassNet(36 * OCTAVES, 12, 2, 7, opt=ake.default_opt(genre=genre))
    net.load_state_dict(golden_weights(genre), strict=True)
    net = net.to(dev).eval()
    est = ake.KeyEstimator(net, SR, FRAMES)
    n_samples = int(round(args.sweep_seconds * SR))
    T = est.frames(n_samples)
    sub = args.sweep_sub
    audio = torch.empty((sub, n_samples), dtype=torch.float32, device=dev)
    synth.synth_batch(rank * sub, sub, n_samples, SR, device=dev, out=audio)
    torch.cuda.synchronize()
    lib = _lib.lib()
    steps, warmup = max(1, min(args.steps, 3)), 1
    for total in [int(x) for x in args.sweep_batches.split(",")]:
        lo, hi = akd.shard_range(total, rank, world)
        n_local = hi - lo
        shard = torch.empty((n_local, akd.ROW), dtype=torch.float32, device=dev)

        def step():
            for b0 in range(0, n_local, sub):
                nb = min(sub, n_local - b0)
                rows, _ = est.estimate_device_rows(audio[:nb])
                shard[b0:b0 + nb].copy_(rows)
            return akd.RowGather(shard, total).wait()

        for _ in range(warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.ake_launch_count(1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            table = step()
        ev1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches = int(lib.ake_launch_count(1))
        assert table.shape == (total, akd.ROW) and bool(torch.isfinite(table).all())
        if rank == 0:
            print(json.dumps({
                "metric": METRIC + " (configs[3] sweep)", "value": total * steps / (ms * 1e-3), "unit": "clips/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
                "dtype": "f32", "data": "synthetic", "gpu_launches": launches,
                "audio_seconds_per_s": total * steps * args.sweep_seconds / (ms * 1e-3),
                "config": {"workload": f"configs[3]: global batch {total} long clips ({args.sweep_seconds:g} s @ {SR} Hz, {n_samples} samples, {T} frames), "
                                       f"{n_local} per GPU in resident sub-batches of {sub} (one synthetic sub-batch re-used), rows all-gathered per step",
                           "global_batch": total, "clips_per_gpu": n_local, "sub_batch": sub, "parallelism": f"dp{world} (batch sharded, logit all-gather)"}}),
                flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """configs[4]: PitchClassNet training step (train_model.py defaults + genre head, batch 8 clips of 151 frames per GPU):
    kept-activation forward + loss + CUDA backward, then one all-reduce of the flat gradient bucket.  With
    --impl reference: the oracle port (float64 torch-CPU autograd, the reference's dtype) on the host cores."""
    import numpy as np
    import torch
    B, T = (args.batch if args.batch != 256 else 8), 1 + SECONDS_STD * SR // (SR // FRAMES)  # --batch N: clips per GPU (default 8, train_model.py)
    rng = np.random.default_rng(0)
    mel = torch.from_numpy(np.log1p(rng.gamma(1.0, 1.0, (B, 1, 36 * OCTAVES, T))).astype(np.float32))
    key_labels = torch.from_numpy((rng.random((B, 12)) < 0.6).astype(np.float32))
    tonic_1h = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, 12, B)), 12)
    genre_1h = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, 11, B)), 11)
    seq = torch.full((B,), T, dtype=torch.int64)
    sd = golden_weights(True)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        from audio_key_estimation_b200 import training
        from oracle import pcn_port
        torch.set_num_threads(os.cpu_count() or 1)
        sd64 = {k: v.double().requires_grad_("running_" not in k) for k, v in sd.items() if v.is_floating_point()}
        def step():
            for v in sd64.values():
                v.grad = None
            out = pcn_port.pcn_forward(sd64, mel.double(), seq, train=True)
            training.criterion(out, key_labels.double(), tonic_1h, genre_1h).backward()
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
        print(json.dumps({"impl": "reference", "metric": "clips/s PitchClassNet training step (fwd+loss+bwd)", "value": B / dt, "unit": "clips/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                          "dtype": "f64", "data": "synthetic", "config": {"workload": f"configs[4]: batch {B} x {T} frames, oracle port autograd, {os.cpu_count()} host threads"},
                          "cpu_baseline": {"value": B / dt, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port", "sample": "whole step"}}), flush=True)
        return
    import torch.distributed as dist
    import audio_key_estimation_b200 as ake
    from audio_key_estimation_b200 import _lib, distributed as akd
    rank, local, world = akd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    net = ake.PitchClassNet(36 * OCTAVES, 12, 2, 7, opt=ake.default_opt(genre=True))
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    ts = ake.TrainStep(net, graph=not args.no_graph)  # the step's ~230 launches replayed as one CUDA graph
    mel, seq, key_labels, tonic_1h, genre_1h = (t.to(dev) for t in (mel, seq, key_labels, tonic_1h, genre_1h))

    def step():
        res = ts.step(mel, seq, key_labels, tonic_1h, genre_1h, assign_grads=False)
        akd.allreduce_gradients(ts.flat_grads)
        return res
    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _lib.lib().ake_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = step()
    ev1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = int(_lib.lib().ake_launch_count(1))
    if ts.use_graph:  # replays do not pass through the library's launch counter: kernels per captured step x steps
        launches = ts.launches_per_step * args.steps
    if rank == 0:
        print(json.dumps({"metric": "clips/s PitchClassNet training step (fwd+loss+bwd+grad all-reduce)", "value": B * world * args.steps / (ms * 1e-3),
                          "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic", "gpu_launches": launches,
                          "loss": float(res["loss"]),
                          "config": {"workload": f"configs[4]: batch {B} clips x {T} frames per GPU, train-mode BN, defaults + genre head, "
                                                 f"{ts.flat_grads.numel()} fp32 gradients all-reduced as one bucket", "parallelism": f"ddp{world}"}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.train:
        run_train(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
