#!/usr/bin/env python
"""bench.py -- denoising steps/s of the LTX-2 DiT and VAE decode frames/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 19b|22b-av|dev-cfg|small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one denoising step of the named configuration: one X0Model forward + the fused denoise update
(CFG guide / Euler step, ltx2_denoise_update).  Weights are a seeded random-init checkpoint of that architecture and
latents/context are synthetic (no network).  Configurations (BASELINE.json `configs`):

  19b      [1]  LTX-2 19B distilled, video-only, 768x512x65 -> N = 3456 tokens, S = 1024, B = 1, 48 blocks  (default)
  22b-av   [2]  LTX-2.3-style audio+video model (cross_attention_adaln, gated attention, a2v/v2a), N = 3456, N_a = 65
  dev-cfg  [3]  LTX-2 19B dev, 1024x768x121 -> N = 12288 tokens, cond+uncond as a batch of 2, CFG 5.0, 25-step schedule
  VAE      [4]  decode_latent at 768x512 x {65,121,241} frames (`vae`, `vae_sweep`)

Keys of the line
  value         steps/s, device-timed (CUDA events) with inputs resident in HBM, max over ranks, EXACTLY --steps steps
  e2e           the same step through the public API (X0Model(Modality(...))) with HOST pinned buffers: H2D of
                latent/context/positions/timesteps and D2H of the denoised sample inside the timed region
  parity        rank 0: the benched model (same weights, same inputs) limited to its first `parity_blocks` blocks + head
                against the CPU oracle on the weights read back from the engine: rel_l2, pearson
                (tests/test_parity.py:53-84 metric); --parity-blocks 48 checks the full model (about a minute of CPU)
  cp_parity     N > 1: the context-parallel forward against the un-sharded forward of the same model on every rank
                (bit-exact with split-K off, max_abs / rel_l2 with the default split-K)
  roofline      the dominant kernel class (tcgen05 GEMM launches): FLOPs / CUDA-event time of those launches inside one
                profiled step, against MEASURED_PEAKS.json bf16_tflops_sustained; attention_frac / step_frac beside it
  cpu_baseline  the oracle (torch-CPU fp32 restatement of the reference; `mlx` is not installable here) on the host
                cores: the parity leg's blocks at the full N, extrapolated to all blocks -- a reported baseline only
  --impl reference  runs only that CPU arm (rank 0) with the same metric/config.

N > 1 ranks: context parallel by default -- ONE sample, token axis sharded over the ranks,
scaling = "strong", value = steps/s of that sample (max over ranks).  `--parallel replicas` runs independent samples
(weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DISTILLED_SIGMAS = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]  # schedulers.py:236-246


def ltx2_scheduler_sigmas(steps, tokens, max_shift=2.05, base_shift=0.95, terminal=0.1):
    """LTX2Scheduler.execute (components/schedulers.py:30-103): shifted, stretched sigma schedule (host floats)."""
    mm = (max_shift - base_shift) / (4096 - 1024)
    shift = math.exp(tokens * mm + (base_shift - mm * 1024))
    lin = [1.0 - i / steps for i in range(steps + 1)]
    sig = [shift / (shift + (1.0 / s - 1.0)) if s != 0 else 0.0 for s in lin]
    scale = (1.0 - sig[steps - 1]) / (1.0 - terminal)
    return [1.0 - (1.0 - s) / scale if s != 0 else 0.0 for s in sig]


CONFIGS = {
    # BASELINE.json configs[1]
    "19b": dict(heads=32, head_dim=128, layers=48, caption=3840, F=9, H=16, W=24, S=1024, B=1, av=False, cfg_scale=1.0,
                name="LTX-2 19B distilled DiT denoise step, 768x512x65 (N=3456 video tokens, S=1024 text tokens, "
                     "B=1, 48 blocks)"),
    # BASELINE.json configs[2]
    "22b-av": dict(heads=32, head_dim=128, layers=48, caption=None, F=9, H=16, W=24, S=1024, B=1, av=True, Na=65,
                   cfg_scale=1.0,
                   name="LTX-2.3-style audio+video DiT denoise step (cross_attention_adaln, gated attention, a2v/v2a), "
                        "768x512x65 (N=3456 video + N_a=65 audio tokens, S=1024, B=1, 48 blocks)"),
    # BASELINE.json configs[3]
    "dev-cfg": dict(heads=32, head_dim=128, layers=48, caption=3840, F=16, H=24, W=32, S=1024, B=2, av=False,
                    cfg_scale=5.0, parity_blocks=1,
                    name="LTX-2 19B dev DiT denoise step, 1024x768x121 (N=12288 video tokens, S=1024), cond+uncond as "
                         "a batch of 2, CFG 5.0, 25-step LTX2Scheduler sigmas, 48 blocks"),
    # CPU-sized debug configuration (not a bench line)
    "small": dict(heads=4, head_dim=128, layers=2, caption=64, F=3, H=4, W=6, S=40, B=1, av=False, cfg_scale=1.0,
                  name="debug 2-block D=512"),
}


def config_sigmas(c):
    if c["cfg_scale"] != 1.0:
        return ltx2_scheduler_sigmas(25, c["F"] * c["H"] * c["W"])
    return DISTILLED_SIGMAS


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tf=float(p["bf16_tflops_sustained"]), tf_burst=float(p["bf16_tflops"]), hbm=float(p["hbm_gbs"]),
                    source="MEASURED_PEAKS.json")
    except Exception:
        return dict(tf=1400.0, tf_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def profiled_traffic(rep_suffix, kernel_substr, skip=0):
    """Mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel from the newest committed
    `ncu --set full` summary (profiles/r2*_kernels.json, else r1c, r1b; tools/summarize_profiles.py); None if absent."""
    try:
        path = next(p for p in (os.path.join(ROOT, "profiles", t + "_kernels.json") for t in ("r2b", "r2", "r1c", "r1b"))
                    if os.path.exists(p))
        with open(path) as f:
            caps = json.load(f)
        rows = [k for name, ks in caps.items() if rep_suffix in name for k in ks if kernel_substr in k["kernel"]][skip:]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in rows:
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(k[m]) * mult.get(k[m + " [unit]"], 1.0)
        return tot / len(rows) if rows else None
    except Exception:
        return None


def flops_per_step(c, cached_text_kv=False):
    """F_blk = 28 N D^2 + 4 S D^2 + 4 N^2 D + 4 N S D; + patchify, caption projection; the
    audio+video model adds the audio stream and the a2v / v2a attentions.  cached_text_kv: the executed count when the
    V1 text K/V (4 S D^2 per block) and the caption projection are reused from the first step of the sample."""
    D = c["heads"] * c["head_dim"]
    N, S, B = c["F"] * c["H"] * c["W"], c["S"], c["B"]
    text_kv = 0 if cached_text_kv else 4 * S * D * D
    blk = 28 * N * D * D + text_kv + 4 * N * N * D + 4 * N * S * D
    extra = 4 * N * 128 * D
    if c["caption"] and not cached_text_kv:
        extra += 2 * S * (c["caption"] * D + D * D)
    if c["av"]:
        Da, Na = 2048, c["Na"]
        blk += 28 * Na * Da * Da + 4 * S * Da * Da + 4 * Na * Na * Da + 4 * Na * S * Da          # audio stream
        blk += 2 * N * D * Da + 4 * Na * Da * Da + 4 * N * Na * Da + 2 * N * Da * D               # a2v
        blk += 2 * Na * Da * Da + 4 * N * D * Da + 4 * Na * N * Da + 2 * Na * Da * Da             # v2a
        blk += 2 * 2 * N * D * 32                                                                 # gate logits
        extra += 4 * Na * 128 * Da
    return B * (c["layers"] * blk + extra)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference block) on the host cores
# ---------------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_block_seconds(c, repeats=1):
    """Time one reference DiT block (oracle) at this config's N, S on all host cores, fp32, one sample."""
    import torch
    from ltx2_b200 import synthetic
    from oracle import dit_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    key = (c["heads"], c["head_dim"], c["F"], c["H"], c["W"], c["S"])
    if key not in _CPU_STATE:
        D = c["heads"] * c["head_dim"]
        cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=1,
                                  cross_attention_dim=D, caption_channels=None)      # context already projected
        w = O.to_engine_keys({k: v for k, v in synthetic.iter_dit_weights(cfg, seed=0)
                              if "transformer_blocks.0." in k})
        N, S = c["F"] * c["H"] * c["W"], c["S"]
        g = torch.Generator().manual_seed(0)
        x = torch.randn(1, N, D, generator=g)
        ctx = torch.randn(1, S, D, generator=g) * 0.1
        ts = torch.randn(1, 1, 6, D, generator=g) * 0.1
        pos = synthetic.video_positions(1, c["F"], c["H"], c["W"])
        pe = O.rope_tables(pos, D, c["heads"], O.MAX_POS)
        _CPU_STATE.clear()
        _CPU_STATE[key] = (w, dict(x=x, context=ctx, timesteps=ts, pe=pe, prompt_timestep=None))
    w, args = _CPU_STATE[key]
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.block(w, 0, args, None, heads=c["heads"], audio_heads=0, v2=False)
            best = min(best, time.perf_counter() - t0)
    return best


def mlx_available():
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    try:
        import mlx.core  # noqa: F401
        return True
    except Exception:
        return False
    finally:
        sys.path.pop(0)


def run_reference(args, c):
    """The CPU arm alone.  `mlx` (the reference's only backend) has no wheel here, so this times the restated oracle.
    Each step is a BOUNDED SAMPLE of a denoising step -- one of the `layers` identical DiT blocks at the full token
    count, one sample of the batch -- and ms_per_step / value are that sample scaled by layers x batch: an
    extrapolation, stated as such (`extrapolated`, `sample_ms`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times = []
    for i in range(args.warmup + args.steps):
        t = cpu_block_seconds(c, 1)
        if i >= args.warmup:
            times.append(t)
    t = sum(times) / len(times)
    v = 1.0 / (t * c["layers"] * c["B"])
    cb = dict(value=v, unit="steps/s", cores=os.cpu_count(), kind="port",
              sample=f"each timed step = 1 of {c['layers']} DiT blocks of the oracle (torch-CPU fp32, {os.cpu_count()} "
                     f"threads) at the full N, one sample: {t * 1e3:.0f} ms; value = 1 / (sample x {c['layers']} blocks "
                     f"x batch {c['B']}), an extrapolation")
    print(json.dumps({
        "impl": "reference", "metric": "denoising steps/sec", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["name"]}, "cpu_baseline": cb, "extrapolated": True, "sample_ms": t * 1e3,
        "sample_fraction_of_step": 1.0 / (c["layers"] * c["B"]),
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mlx_available": mlx_available(),
    }))
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU arm: VAE
# ---------------------------------------------------------------------------------------------------------
def vae_conv_flops(T, H, W, blocks=None, base=128):
    """sum over convs of 2*Cin*Cout*27*T*H*W for one SimpleVideoDecoder pass."""
    from ltx2_b200 import synthetic
    cfg = synthetic.VaeConfig(decoder_blocks=blocks or synthetic.DEFAULT_DECODER_BLOCKS, base_channels=base)
    total, C = 2 * 128 * (base * 8) * 27 * T * H * W, base * 8
    for kind, p, c in cfg.stages():
        if kind == "res":
            total += p["num_layers"] * 2 * (2 * c * c * 27 * T * H * W)
        else:
            s = p["stride"]
            total += 2 * c * (s[0] * s[1] * s[2] * c // p["multiplier"]) * 27 * T * H * W
            T, H, W, C = T * s[0] - (1 if s[0] > 1 else 0), H * s[1], W * s[2], c // p["multiplier"]
    return total + 2 * C * 48 * 27 * T * H * W


def pearson(a, b):
    import numpy as np
    return float(np.corrcoef(a.double().flatten().cpu().numpy(), b.double().flatten().cpu().numpy())[0, 1])


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def bench_vae(args, dev, rank, world=1):
    """decode_latent of 9/16/31-frame latents (65/121/241 frames @ 512x768) exactly as the reference schedules it
    (7-frame chunks, overlap 2, cross-fade, uint8), device-timed; the conv kernel's roofline; parity of the benched
    decoder against the oracle.  world > 1: the decode units are spread over the ranks and collected on rank 0
    (strong scaling)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ltx2_b200 import _lib, synthetic
    from ltx2_b200.video_vae import (SimpleVideoDecoder, chunk_plan, decode_latent, decode_latent_video,
                                     disable_temporal_shards, enable_temporal_shards)
    vcfg = synthetic.VaeConfig()
    dec = SimpleVideoDecoder(device=dev)
    w_cpu = {}
    want_parity = rank == 0 and not args.no_parity

    def tee():
        for k, t in synthetic.iter_vae_weights(vcfg, seed=0, device=dev, dtype=torch.bfloat16):
            if want_parity:
                w_cpu[k] = t.float().cpu()
            yield k, t

    dec.load_weights(tee())
    assert not dec.missing_weights()
    kw = dict(group=dist.group.WORLD, dst=0) if world > 1 else {}
    shard_parity = None
    if world > 1:
        # bit-exactness of the sharded decode against this rank's own single-GPU decode of the same clip (noise off)
        plat = synthetic.latents((1, 128, 7, 16, 24), seed=45).to(dev)
        keep = dec.decode_noise_scale
        dec.decode_noise_scale = 0.0
        ref = decode_latent_video(plat, dec)
        enable_temporal_shards(dec, (1, 128, 7, 16, 24), group=dist.group.WORLD)
        got = decode_latent_video(plat, dec, group=dist.group.WORLD)
        dec.decode_noise_scale = keep
        t = torch.tensor([0.0 if torch.equal(got, ref) else 1.0, float((got - ref).abs().max())], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shard_parity = {"bit_exact": float(t[0]) == 0.0, "max_abs": float(t[1]),
                        "what": f"49-frame chunk (latent 1x128x7x16x24) decoded in temporal shards over {world} ranks vs the "
                                f"single-GPU decode of the same latent on every rank, noise off"}
        del ref, got

        def time_chunk(t_lat):
            from ltx2_b200.video_vae import decode_sharded
            cl = synthetic.latents((1, 128, t_lat, 16, 24), seed=46).to(dev)
            for _ in range(2):
                decode_sharded(dec, cl, 0.05, dst=0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            a.record()
            for _ in range(6):
                decode_sharded(dec, cl, 0.05, dst=0)
            b.record()
            torch.cuda.synchronize()
            tt = torch.tensor([a.elapsed_time(b) / 6], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt[0])

        shard_parity["chunk_ms"] = {"7_latent_frames": time_chunk(7), "4_latent_frames": time_chunk(4),
                                    "note": "one sharded SimpleVideoDecoder call incl. assembling the clip on rank 0"}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_decode(t_lat, n):
        lat = synthetic.latents((1, 128, t_lat, 16, 24), seed=43).to(dev)
        for _ in range(3):
            out = decode_latent(lat, dec, **kw)
        l0 = _lib.lib().ltx2_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        for _ in range(n):
            out = decode_latent(lat, dec, **kw)
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / n
        launches = (_lib.lib().ltx2_launch_count() - l0) // n
        frames = 8 * (t_lat - 1) + 1
        assert rank != 0 or out.shape == (frames, 512, 768, 3)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            t2 = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
            dist.all_reduce(t2, op=dist.ReduceOp.SUM)
            launches = int(t2[0])
        return lat, frames, ms, launches

    n = max(3, min(args.steps, 8))
    lat, frames, ms, launches = time_decode(9, n)
    # e2e: host latent in, uint8 frames on the host out
    lat_h = lat.cpu().pin_memory()
    out_h = torch.empty(frames, 512, 768, 3, dtype=torch.uint8).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(n):
        r = decode_latent(lat_h, dec, **kw)
        if r is not None:
            out_h.copy_(r, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    sync()
    ms_e2e = e0.elapsed_time(e1) / n
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t[0])
    # roofline of the conv kernel: one profiled pass over the first (7 latent frame) chunk
    L = _lib.lib()
    _lib.check(L.ltx2_vae_set_profile(dec._h, 1))
    dec(lat[:, :, :7].contiguous(), timestep=0.05)
    pm, pf, pl = C.c_double(), C.c_double(), C.c_int64()
    _lib.check(L.ltx2_vae_profile_read(dec._h, C.byref(pm), C.byref(pf), C.byref(pl)))
    _lib.check(L.ltx2_vae_set_profile(dec._h, 0))
    pk = peaks()
    tf = pf.value / (pm.value * 1e-3) / 1e12 if pm.value > 0 else 0.0
    plan = chunk_plan(9)
    alg = sum(vae_conv_flops(b - a, 16, 24) for a, b in plan)
    # sweep (BASELINE.json configs[4])
    sweep = [{"frames": frames, "latent_frames": 9, "ms_per_decode": ms, "frames_per_s": frames * 1000.0 / ms}]
    if not args.no_sweep:
        for t_lat in (16, 31):
            _, fr, ms_s, _ = time_decode(t_lat, max(2, n // 2))
            alg_s = sum(vae_conv_flops(b - a, 16, 24) for a, b in chunk_plan(t_lat))
            sweep.append({"frames": fr, "latent_frames": t_lat, "ms_per_decode": ms_s, "frames_per_s": fr * 1000.0 / ms_s,
                          "decode_frac": alg_s / (ms_s * 1e-3) / 1e12 / (pk["tf"] * world)})
    sweep[0]["decode_frac"] = alg / (ms * 1e-3) / 1e12 / (pk["tf"] * world)
    # parity of the benched decoder (same weights) against the oracle on a 2-frame latent of the benched H x W
    parity = None
    if want_parity:
        from oracle import vae_oracle as V
        torch.set_num_threads(os.cpu_count() or 1)
        plat = synthetic.latents((1, 128, 2, 16, 24), seed=44)
        keep = dec.decode_noise_scale
        dec.decode_noise_scale = 0.0
        got = dec(plat, timestep=0.05).float().cpu()
        dec.decode_noise_scale = keep
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = V.vae_decode(w_cpu, plat, decoder_blocks=synthetic.DEFAULT_DECODER_BLOCKS, base_channels=128,
                               timestep=0.05)
        t_or = time.perf_counter() - t0
        parity = {"rel_l2": rel_l2(got, ref), "pearson": pearson(got, ref), "max_abs": float((got - ref).abs().max()),
                  "what": "benched decoder (same weights) on a 1x128x2x16x24 latent -> 9 frames @ 512x768, noise off, vs "
                          "oracle.vae_oracle.vae_decode (fp32 CPU)", "tolerance": {"rel_l2": 3e-2, "pearson": 0.999},
                  "oracle_seconds": t_or,
                  "cpu_frames_per_s": 9.0 / t_or}
        parity["ok"] = bool(parity["rel_l2"] < 3e-2 and parity["pearson"] > 0.999)
    if world > 1:
        disable_temporal_shards(dec)
    return {
        "metric": "VAE decode frames/sec", "value": frames * 1000.0 / ms, "unit": "frames/s", "ms_per_decode": ms,
        "config": {"workload": "decode_latent, latent 1x128x9x16x24 -> 65 frames @ 512x768, V2.0 decoder stack "
                               "(base 128, 5 res blocks/group), reference chunking 7/2 -> chunks " + str(plan),
                   "parallelism": "single GPU" if world == 1 else
                                  f"temporal shards: the frames of every chunk are split over {world} ranks (one halo frame "
                                  f"per neighbour and conv through peer memory), collected on rank 0 (strong scaling)",
                   "noise": "decode_noise_scale 0.025 (reference default), timestep 0.05"},
        "gpu_launches": int(launches),
        "e2e": {"value": frames * 1000.0 / ms_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(lat_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel())},
        "parity": parity, "shard_parity": shard_parity, "sweep": sweep,
        "roofline": {"bound": "tensor", "kernel": "conv3d_kernel (tcgen05 implicit GEMM, all convs of one 7-frame chunk)",
                     "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"],
                     "traffic": profiled_traffic("prof_conv", "conv3d_kernel<128>"),
                     "traffic_note": "DRAM bytes of one last-stage conv launch (128->128 ch, 49x128x192) from the newest "
                                     "profiles/*_kernels.json; algorithmic bytes 330 MB in + 308 MB out + 0.9 MB weights",
                     "launches": int(pl.value), "ms_in_decode": pm.value, "flops_in_decode": pf.value,
                     "decode": {"algorithmic_flops": alg, "achieved": alg / (ms * 1e-3) / 1e12,
                                "frac": alg / (ms * 1e-3) / 1e12 / (pk["tf"] * world)}},
    }


# ---------------------------------------------------------------------------------------------------------
# GPU arm: DiT
# ---------------------------------------------------------------------------------------------------------
class DitBench:
    """Model + synthetic inputs + the step function of one configuration."""

    def __init__(self, c, dev, rank, world, cp, fp8=False):
        import torch
        from ltx2_b200 import synthetic
        from ltx2_b200.loader import iter_engine_weights
        from ltx2_b200.transformer import LTXModel, LTXModelType, X0Model
        self.c, self.dev, self.rank, self.world, self.cp = c, dev, rank, world, cp
        D = c["heads"] * c["head_dim"]
        av = c["av"]
        self.cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"],
                                       num_layers=c["layers"], cross_attention_dim=D, caption_channels=c["caption"],
                                       cross_attention_adaln=av, apply_gated_attention=av, audio=av)
        self.model = LTXModel(model_type=LTXModelType.AudioVideo if av else LTXModelType.VideoOnly,
                              num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=c["layers"],
                              cross_attention_dim=D, caption_channels=c["caption"], cross_attention_adaln=av,
                              apply_gated_attention=av, av_ca_timestep_scale_multiplier=1000, device=dev,
                              fp8_linear=fp8)
        self.model.load_weights(iter_engine_weights(
            synthetic.iter_dit_weights(self.cfg, seed=0, device=dev, dtype=torch.bfloat16), include_audio=av))
        assert not self.model.missing_weights()
        self.x0model = X0Model(self.model)
        self.N, self.S, self.B = c["F"] * c["H"] * c["W"], c["S"], c["B"]
        self.sigmas = config_sigmas(c)
        self.n_sig = len(self.sigmas) - 1
        ctx_ch = c["caption"] or D
        B = self.B
        # one sample; a CFG batch holds the same latent twice (cond | uncond) with two contexts
        self.lat0 = synthetic.latents((1, self.N, 128), seed=42 + (0 if cp else rank))
        self.ctx0 = synthetic.latents((B, self.S, ctx_ch), seed=7, std=0.1).to(torch.bfloat16)
        self.pos0 = synthetic.video_positions(B, c["F"], c["H"], c["W"], fps=24.0)
        self.lat_d, self.ctx_d, self.pos_d = self.lat0.to(dev), self.ctx0.to(dev), self.pos0.to(dev)
        self.sig_d = [torch.full((B,), s, device=dev) for s in self.sigmas]
        if av:
            Na = c["Na"]
            self.alat0 = synthetic.latents((1, Na, 128), seed=142)
            self.actx0 = synthetic.latents((B, self.S, 2048), seed=8, std=0.1).to(torch.bfloat16)
            self.apos0 = synthetic.audio_positions(B, Na)
            self.alat_d, self.actx_d, self.apos_d = self.alat0.to(dev), self.actx0.to(dev), self.apos0.to(dev)

    def enable_cp(self):
        from ltx2_b200 import context_parallel
        context_parallel.enable(self.model, batch=self.B, n_total=self.N, context_tokens=self.S)

    def modalities(self, i, latent, alatent=None, host=False):
        from ltx2_b200.transformer import Modality
        k = i % self.n_sig
        sig = self.sig_h[k] if host else self.sig_d[k]
        xin = latent if self.B == 1 else latent.expand(self.B, -1, -1).contiguous()
        v = Modality(latent=xin, context=self.ctx_h if host else self.ctx_d, context_mask=None, timesteps=sig,
                     positions=self.pos_h if host else self.pos_d, sigma=sig)
        a = None
        if self.c["av"]:
            a = Modality(latent=alatent, context=self.actx_h if host else self.actx_d, context_mask=None, timesteps=sig,
                         positions=self.apos_h if host else self.apos_d, sigma=sig)
        return v, a

    def step_device(self, i, latent, alatent=None):
        """One denoising step with everything resident in HBM: X0 forward + the fused update kernel."""
        from ltx2_b200 import sampling
        k = i % self.n_sig
        s, s_next = self.sigmas[k], self.sigmas[k + 1]
        v, a = self.modalities(i, latent, alatent)
        out = self.x0model(v, a) if a is not None else self.x0model(v)
        if a is not None:
            x0v, x0a = out
            return (sampling.denoise_update(latent, x0v, s, s_next), sampling.denoise_update(alatent, x0a, s, s_next))
        if self.B == 2:     # cond | uncond -> CFGGuider.guide + Euler step in one kernel
            return sampling.denoise_update(latent, out[:1], s, s_next, uncond_x0=out[1:], cfg_scale=self.c["cfg_scale"]), None
        return sampling.denoise_update(latent, out, s, s_next), None

    def pin_host(self):
        import torch
        self.lat_h, self.ctx_h, self.pos_h = self.lat0.pin_memory(), self.ctx0.pin_memory(), self.pos0.pin_memory()
        self.sig_h = [torch.full((self.B,), s).pin_memory() for s in self.sigmas]
        h2d = self.B * self.lat_h.numel() * 4 + self.ctx_h.numel() * 2 + self.pos_h.numel() * 4 + 4 * self.B
        d2h = self.B * self.lat_h.numel() * 4
        if self.c["av"]:
            self.alat_h, self.actx_h, self.apos_h = self.alat0.pin_memory(), self.actx0.pin_memory(), self.apos0.pin_memory()
            h2d += self.alat_h.numel() * 4 + self.actx_h.numel() * 2 + self.apos_h.numel() * 4 + 4 * self.B
            d2h += self.alat_h.numel() * 4
        self.out_h = torch.empty(self.B, self.N, 128).pin_memory()
        self.aout_h = torch.empty(1, self.c.get("Na", 1), 128).pin_memory()
        return h2d, d2h

    def step_e2e(self, i):
        import torch
        if i % self.n_sig == 0:
            self.model.reset_context_cache()      # a new sample (see run_steps)
        v, a = self.modalities(i, self.lat_h, self.alat_h if self.c["av"] else None, host=True)
        out = self.x0model(v, a) if a is not None else self.x0model(v)
        if a is not None:
            self.out_h.copy_(out[0], non_blocking=True)
            self.aout_h.copy_(out[1], non_blocking=True)
        else:
            self.out_h.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the result (mx.eval analogue)

    # ---- parity of the benched model against the oracle ----
    def parity(self, blocks):
        """The benched model limited to its first `blocks` blocks (+ head), same weights and inputs, against the oracle
        on the weights read back from the engine (bf16 values widened to fp32).  Sample 0 of the batch."""
        import torch
        from oracle import dit_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        c = self.c
        blocks = min(blocks, c["layers"])
        m = self.model
        w = {}
        for k in m.weight_keys():
            if k.startswith("transformer_blocks."):
                if int(k.split(".")[1]) >= blocks:
                    continue
            w[k] = m.get_weight(k).cpu()
        m.set_layer_limit(blocks)
        sig = self.sigmas[5 % self.n_sig]
        ts = torch.full((self.B,), sig, device=self.dev)
        from ltx2_b200.transformer import Modality
        lat = self.lat_d.expand(self.B, -1, -1).contiguous()
        vmod = Modality(latent=lat, context=self.ctx_d, context_mask=None, timesteps=ts, positions=self.pos_d, sigma=ts)
        amod = None
        if c["av"]:
            amod = Modality(latent=self.alat_d, context=self.actx_d, context_mask=None, timesteps=ts,
                            positions=self.apos_d, sigma=ts)
        out = self.x0model(vmod, amod) if amod is not None else self.x0model(vmod)
        m.set_layer_limit(0)
        got_v = (out[0] if amod is not None else out)[:1].float().cpu()
        got_a = out[1].float().cpu() if amod is not None else None
        t1 = torch.tensor([sig])
        vd = dict(latent=self.lat0, context=self.ctx0[:1].float(), timesteps=t1, positions=self.pos0[:1], sigma=t1)
        ad = None
        kw = dict(num_layers=blocks, heads=c["heads"])
        if c["av"]:
            ad = dict(latent=self.alat0, context=self.actx0[:1].float(), timesteps=t1, positions=self.apos0[:1], sigma=t1)
            kw.update(audio_heads=32, v2=True, av_ca_timestep_scale_multiplier=1000)
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = O.x0_forward(w, vd, ad, **kw)
        t_or = time.perf_counter() - t0
        ref_v = ref[0] if ad is not None else ref
        res = {"rel_l2": rel_l2(got_v, ref_v), "pearson": pearson(got_v, ref_v),
               "max_abs": float((got_v - ref_v).abs().max()), "blocks": blocks, "of_blocks": c["layers"],
               "what": f"x0 of the benched model (same weights, sample 0, sigma {sig}) limited to its first {blocks} of "
                       f"{c['layers']} blocks + output head vs oracle.dit_oracle (fp32 CPU on the bf16 weights read back "
                       f"from the engine)", "tolerance": {"rel_l2": 2e-2, "pearson": 0.999}, "oracle_seconds": t_or}
        if ad is not None:
            res["audio_rel_l2"] = rel_l2(got_a, ref[1])
            res["audio_pearson"] = pearson(got_a, ref[1])
        res["ok"] = bool(res["rel_l2"] < 2e-2 and res["pearson"] > 0.999)
        return res


def run_dit(args, c, dev, rank, local_rank, world, name, fp8=False):
    """Time one DiT configuration; returns the result dict on rank 0 (None elsewhere).  fp8: the engine's FP8 linear
    path (E4M3 weights + per-token E4M3 activations for the norm-fed linears) on the same synthetic checkpoint."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ltx2_b200 import _lib, context_parallel

    cp = world > 1 and args.parallel == "cp"
    b = DitBench(c, dev, rank, world, cp, fp8=fp8)
    model = b.model

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity against the oracle (un-sharded model, rank 0) ----
    parity = None
    if rank == 0 and not args.no_parity:
        parity = b.parity(min(args.parity_blocks, c.get("parity_blocks", args.parity_blocks)))
    barrier()

    # ---- context parallel: sharded forward vs the un-sharded forward of the same model, on every rank ----
    cp_parity = None
    alat = b.alat_d if c["av"] else None
    if cp:
        # Two comparisons.  (1) EXACT: with the kernel choice pinned (standard GEMM tiles, every attention query tile as
        # a split-KV item -- a pair item and a split-KV item round differently and which tiles pair up depends on the
        # local token count --, no split-K) every output element is computed by the same instruction sequence on one GPU and on P GPUs, so the
        # sharded forward must be BIT-IDENTICAL -- this checks the sharding, the head exchange and the context broadcast.
        # (2) DEFAULT: shard-shaped GEMM kernels, makespan-optimal attention items and split-K change fp32 summation
        # orders; the bound is rel_l2 2e-3, a tenth of the engine-vs-oracle tolerance (FP8 linears: 1e-2 -- a last-bit
        # difference before the per-token E4M3 quantisation moves a whole quantisation step after it).
        # LTX2_ATTN_SPLIT=0: the SM-pair attention kernel (>= 8192 tokens) with whole items per cluster -- its stream-K
        # cut points depend on the number of (batch, head) slices, i.e. on the rank count
        canon = {"LTX2_GEMM_T": "0", "LTX2_GEMM_2CTA": "0", "LTX2_ATTN_PAIRS": "0", "LTX2_ATTN_SPLIT": "0"}
        v, a = b.modalities(5, b.lat_d, alat)

        def fwd(env=None):
            old = {k: os.environ.get(k) for k in (env or {})}
            os.environ.update(env or {})
            try:
                model.reset_context_cache()
                out = b.x0model(v, a) if a is not None else b.x0model(v)
                out = [t.clone() for t in (out if isinstance(out, tuple) else (out,))]
                torch.cuda.synchronize()
            finally:
                for k, val in old.items():
                    if val is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = val
            return out

        ref_c, ref_d = fwd(canon), fwd()
        b.enable_cp()
        context_parallel.set_split_k(model, 1)
        out_c = fwd(canon)
        context_parallel.set_split_k(model, 0)
        out_d = fwd()
        exact = all(torch.equal(o, r) for o, r in zip(out_c, ref_c))
        mx_c = max(float((o - r).abs().max()) for o, r in zip(out_c, ref_c))
        mx = max(float((o - r).abs().max()) for o, r in zip(out_d, ref_d))
        rl = max(float((o - r).norm() / r.norm()) for o, r in zip(out_d, ref_d))
        t = torch.tensor([0.0 if exact else 1.0, mx_c, mx, rl], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cp_parity = {"bit_exact_pinned_kernels": float(t[0]) == 0.0, "max_abs_pinned_kernels": float(t[1]),
                     "max_abs": float(t[2]), "rel_l2": float(t[3]),
                     "what": f"x0 of the full {c['layers']}-block model: context-parallel forward over {world} ranks vs "
                             f"the un-sharded forward of the same model, max over ranks.  pinned kernels (standard GEMM "
                             f"tiles, split-KV attention items only, split-K off): must be bit-exact; default kernel choice "
                             f"(shard-shaped GEMMs, split-K): rel_l2 <= {1e-2 if fp8 else 2e-3:g}",
                     "ok": bool(float(t[0]) == 0.0 and float(t[3]) <= (1e-2 if fp8 else 2e-3))}

    def run_steps(n, first=0):
        latent, al = b.lat_d.clone(), (b.alat_d.clone() if c["av"] else None)
        for i in range(first, first + n):
            if i % b.n_sig == 0:
                model.reset_context_cache()       # a new sample: its first step projects the text K/V again
            latent, al = b.step_device(i, latent, al)
        return latent

    run_steps(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = _lib.lib().ltx2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    latent = run_steps(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib().ltx2_launch_count() - launches0
    clocks = sampler.stop()
    finite = bool(torch.isfinite(latent).all())

    # ---- a longer region at N > 1 (the K-step region is a fraction of a second there): >= 100 steps, clocks sampled ----
    long_run = None
    if world > 1 and args.steps < 100 and not args.no_long and name == args.config:
        n_long = 100
        barrier()
        sampler = ClockSampler(local_rank)
        e0.record()
        run_steps(n_long)
        e1.record()
        barrier()
        ms_long = e0.elapsed_time(e1)
        t = torch.tensor([ms_long], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        long_run = {"steps": n_long, "ms_per_step": float(t[0]) / n_long, "clocks": sampler.stop()}
        long_run["value"] = (1 if cp else world) * 1000.0 / long_run["ms_per_step"]

    # ---- e2e: host buffers in, host result out, every step ----
    h2d, d2h = b.pin_host()
    for i in range(min(args.warmup, 3)):
        b.step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        b.step_e2e(i)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)

    # ---- roofline of the dominant kernel class, from one profiled step ----
    L = _lib.lib()
    _lib.check(L.ltx2_dit_set_profile(model._h, 1))
    b.step_device(1, b.lat_d.clone(), b.alat_d.clone() if c["av"] else None)
    pm, pf, pl = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int64 * 3)()
    _lib.check(L.ltx2_dit_profile_read(model._h, pm, pf, pl, 3))
    _lib.check(L.ltx2_dit_set_profile(model._h, 0))

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if cp:
        context_parallel.disable(model)
    b_n_sig = b.n_sig
    del b, model
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    pk = peaks()
    ms_step = ms / args.steps
    samples = 1 if cp else world
    value = samples * 1000.0 / ms_step
    gemm_tf = pf[0] / (pm[0] * 1e-3) / 1e12 if pm[0] > 0 else 0.0
    attn_tf = pf[1] / (pm[1] * 1e-3) / 1e12 if pm[1] > 0 else 0.0
    text_cached = not c["av"]
    fl_ref = flops_per_step(c)
    # one step in n_sig (the first of every sample) recomputes the text K/V
    fl_exec = (flops_per_step(c, cached_text_kv=text_cached) * (b_n_sig - 1) + fl_ref) / b_n_sig
    step_tf = samples * fl_exec / (ms_step * 1e-3) / 1e12
    out = {
        "metric": "denoising steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if cp else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": c["name"], "name": name,
                   "parallelism": (f"cp{world}: token axis sharded, heads re-sharded for self-attention by peer-memory "
                                   f"stores fused into the q/k-norm and attention kernels" if cp else
                                   (f"replicas x{world}" if world > 1 else "single")),
                   "l2": "weights read per step (25.8 GB bf16) exceed the 126 MB L2; no flush needed",
                   "weights": "seeded random init, reference key names", "residual_stream": "fp32",
                   "gemm_operands": "bf16, fp32 accumulate",
                   "text_kv": ("V1 text K/V + caption projection computed on the first step of a sample and reused "
                               "(context is sigma-independent, transformer.py:427-455)" if text_cached else
                               "recomputed every step (cross_attention_adaln: K/V depend on sigma)")},
        "clocks": clocks, "gpu_launches": int(launches), "finite": finite,
        "parity": parity, "cp_parity": cp_parity,
        "e2e": {"value": samples * 1000.0 * args.steps / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gemm_frac": gemm_tf / pk["tf"], "attention_frac": attn_tf / pk["tf"], "step_frac": step_tf / (pk["tf"] * world),
        "flops_per_step": {"as_reference": fl_ref, "executed": fl_exec,
                           "note": "step_frac uses the EXECUTED count: V1 text K/V + caption projection are computed on the "
                                   "first step of every sample (the timed loop restarts a sample every n_sigma steps) and "
                                   "reused on the others"},
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all DiT linears of one step)",
                     "achieved": gemm_tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf"],
                     "traffic": profiled_traffic("prof_gemm", "gemm_bf16_kernel", skip=3),
                     "traffic_note": "mean DRAM bytes per launch over the per-block GEMMs of one block from the newest "
                                     "profiles/*_kernels.json; algorithmic bytes of those launches average 128 MB",
                     "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                     "launches": int(pl[0]), "ms_in_step": pm[0], "flops_in_step": pf[0],
                     "attention": {"kernel": "attention_pair_kernel (tcgen05, two softmax streams per CTA; head_dim 128)",
                                   "achieved": attn_tf, "frac": attn_tf / pk["tf"], "launches": int(pl[1]),
                                   "ms_in_step": pm[1], "flops_in_step": pf[1],
                                   "traffic": profiled_traffic("prof_attn", "attention_pair_kernel"),
                                   "traffic_note": "mean DRAM bytes per launch over one self- and one text "
                                                   "cross-attention launch; algorithmic bytes: Q, K, V, O of the launch "
                                                   "= 113 MB / 50 MB"},
                     "step": {"algorithmic_flops": fl_exec, "achieved": step_tf, "frac": step_tf / (pk["tf"] * world),
                              "note": "executed whole-step FLOPs over all ranks / step time, against world x the "
                                      "per-GPU peak"}},
    }
    if fp8:
        g8 = pf[2] / (pm[2] * 1e-3) / 1e12 if pm[2] > 0 else 0.0
        out["dtype"] = "fp8-e4m3 x fp8-e4m3 (self-attn QKV, text-attn Q, FFN up + down) + bf16 (attention, out-projections, rest)"
        out["config"]["gemm_operands"] = ("E4M3 weights (one scale per output row) x E4M3 activations (dynamic absmax scale "
                                          "per token from the norm kernel; FFN hidden against a per-token Cauchy-Schwarz "
                                          "bound) on tcgen05.mma kind::f8f6f4 for QKV / text-Q / FFN up / FFN down; bf16 "
                                          "for the attention out-projections, text K/V, head; fp32 accumulate")
        out["parity"]["tolerance"] = {"rel_l2": 6e-2, "pearson": 0.995}
        out["parity"]["ok"] = bool(out["parity"]["rel_l2"] < 6e-2 and out["parity"]["pearson"] > 0.995)
        out["parity"]["what"] += ("; the oracle runs on the DEQUANTISED E4M3 weights read back from the engine, so the "
                                  "difference is the E4M3 rounding of the activations (3 mantissa bits)")
        out["fp8_gemm"] = {"kernel": "gemm_bf16_kernel<BN, FP8=true> (tcgen05.mma kind::f8f6f4)", "achieved": g8,
                           "unit": "TFLOP/s", "launches": int(pl[2]), "ms_in_step": pm[2], "flops_in_step": pf[2],
                           "frac_of_2x_bf16_sustained": g8 / (2 * pk["tf"]),
                           "note": "no measured FP8 peak in MEASURED_PEAKS.json: the denominator is twice the measured "
                                   "bf16 sustained rate (the nominal FP8:bf16 ratio)"}
    if long_run is not None:
        out["long_run"] = long_run
        if (clocks.get("samples") or 0) < 5:
            out["clocks"] = long_run["clocks"]
            out["clocks"]["from"] = "long_run (100 steps): the K-step region is too short for the 100 ms sampler"
    return out


def run_ours(args, c):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.vae_only:
        vae = bench_vae(args, dev, rank, world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            vae["n_gpus"] = world
            print(json.dumps(vae))
        return 0
    out = run_dit(args, c, dev, rank, local_rank, world, args.config)

    # ---- the other BASELINE.json configurations, attached as `configs` (each a full line of its own) ----
    extra = [x for x in (args.extra_configs.split(",") if args.extra_configs else []) if x]
    if args.extra_configs is None and world == 8 and args.config == "19b":
        extra = ["22b-av", "dev-cfg"]
    extras = []
    for name in extra:
        try:
            r = run_dit(args, CONFIGS[name], dev, rank, local_rank, world, name)
        except Exception as e:      # an extra configuration must not take the headline line down with it
            r = {"config": {"name": name}, "error": f"{type(e).__name__}: {e}"[:400]}
            if world > 1:
                raise
        if rank == 0:
            extras.append(r)

    # ---- the FP8 linear path on the same configuration, reported BESIDE the bf16 line (never instead of it) ----
    fp8_line = None
    if args.fp8 and (world == 1 or args.parallel == "cp"):
        fp8_line = run_dit(args, c, dev, rank, local_rank, world, args.config, fp8=True)

    # ---- second half of the metric: VAE decode frames/s ----
    vae = None
    if args.config == "19b" and not args.no_vae:
        vae = bench_vae(args, dev, rank, world)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    if extras:
        out["configs"] = extras
    if fp8_line is not None:
        out["fp8"] = fp8_line
    if vae is not None:
        out["vae"] = vae
        out["vae_frames_per_s"] = vae["value"]
        out["vae_decode_frac"] = vae["roofline"]["decode"]["frac"]
        out["vae_sweep"] = vae["sweep"]
    if world == 1 and not args.no_cpu:
        p = out.get("parity")
        if p is not None:
            # the parity leg already ran the oracle at the full token count: reuse its clock
            t_blk = p["oracle_seconds"] / p["blocks"]
            out["cpu_baseline"] = dict(
                value=1.0 / (t_blk * c["layers"] * c["B"]), unit="steps/s", cores=os.cpu_count(), kind="port",
                seconds_per_block=t_blk,
                sample=f"{p['blocks']} of {c['layers']} DiT blocks + prepare + output head of the oracle (torch-CPU fp32, "
                       f"{os.cpu_count()} threads) at the full N, one sample: {p['oracle_seconds']:.1f} s, extrapolated to "
                       f"{c['layers']} blocks x batch {c['B']}; mlx (the reference's backend) is not installable here, so "
                       f"this is the restated oracle, not MLX")
        else:
            t = cpu_block_seconds(c)
            out["cpu_baseline"] = dict(value=1.0 / (t * c["layers"] * c["B"]), unit="steps/s", cores=os.cpu_count(),
                                       kind="port", seconds_per_block=t,
                                       sample=f"1 of {c['layers']} DiT blocks (oracle, torch-CPU fp32) at the full N, "
                                              f"extrapolated")
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="19b", choices=list(CONFIGS))
    ap.add_argument("--extra-configs", default=None,
                    help="comma list of further configurations to run and attach as `configs` (default: 22b-av,dev-cfg "
                         "when --gpus 8 --config 19b; '' = none)")
    ap.add_argument("--parity-blocks", type=int, default=4,
                    help="blocks of the benched model checked against the CPU oracle (48 = the full model, ~1 min)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity legs")
    ap.add_argument("--fp8", action=argparse.BooleanOptionalAction, default=True,
                    help="also time the FP8 linear path of the same configuration (`fp8` key; at N > 1 under context parallelism)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-vae", action="store_true", help="skip the VAE decode leg")
    ap.add_argument("--vae-only", action="store_true", help="only the VAE decode leg (prints its object as the line)")
    ap.add_argument("--no-sweep", action="store_true", help="VAE: only the 65-frame point")
    ap.add_argument("--no-long", action="store_true", help="N > 1: skip the extra 100-step region")
    ap.add_argument("--parallel", default="cp", choices=["cp", "replicas"],
                    help="N>1: context-parallel single sample (strong scaling) or independent replicas (weak)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    c = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, c)
    return run_ours(args, c)


if __name__ == "__main__":
    sys.exit(main())
