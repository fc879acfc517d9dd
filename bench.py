#!/usr/bin/env python
"""bench.py -- denoising steps/s of the LTX-2 19B DiT on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 19b|small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one denoising step of BASELINE.json configs[1] (LTX-2 19B distilled, 768x512x65 -> N = 3456
video tokens, S = 1024 text tokens, B = 1, 48 blocks): one X0Model forward + the Euler update.  Weights are a
seeded random-init checkpoint of that architecture and latents/context are synthetic (no network).

  value     steps/s, device-timed (CUDA events) with inputs resident in HBM, max over ranks
  e2e       the same step through the public API (X0Model(Modality(...)) with HOST pinned buffers: H2D of
            latent/context/positions/timesteps and D2H of the denoised sample inside the timed region
  roofline  the dominant kernel class (tcgen05 GEMM launches): algorithmic FLOPs / CUDA-event time of those
            launches inside one profiled step, against MEASURED_PEAKS.json bf16_tflops_sustained
  cpu_baseline  the oracle (torch-CPU fp32 restatement of the reference block; `mlx` is not installable here)
            on the host cores: one of the 48 blocks at the same N, extrapolated x48 -- a reported baseline only
  --impl reference  runs only that CPU arm (rank 0) with the same metric/config.

N > 1 ranks: context parallel by default -- ONE sample, token axis sharded over the ranks,
scaling = "strong", value = steps/s of that sample (max over ranks).  `--parallel replicas` runs independent samples
(weak scaling, no data-path collective).  The VAE leg at N > 1 decodes on rank 0 only (its two temporal chunks do
not fill more GPUs at 65 frames; multi-GPU VAE is next-round work).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DISTILLED_SIGMAS = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]  # schedulers.py:236-246

CONFIGS = {
    # BASELINE.json configs[1]
    "19b": dict(heads=32, head_dim=128, layers=48, caption=3840, F=9, H=16, W=24, S=1024,
                name="LTX-2 19B distilled DiT denoise step, 768x512x65 (N=3456 video tokens, S=1024 text tokens, "
                     "B=1, 48 blocks)"),
    # CPU-sized debug configuration (not a bench line)
    "small": dict(heads=4, head_dim=128, layers=2, caption=64, F=3, H=4, W=6, S=40, name="debug 2-block D=512"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tf=float(p["bf16_tflops_sustained"]), tf_burst=float(p["bf16_tflops"]), hbm=float(p["hbm_gbs"]),
                    source="MEASURED_PEAKS.json")
    except Exception:
        return dict(tf=1400.0, tf_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def profiled_traffic(rep_suffix, kernel_substr, skip=0):
    """Mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel from the committed
    `ncu --set full` summary (profiles/r1c_kernels.json, else r1b; produced by tools/summarize_profiles.py); None if
    absent."""
    try:
        path = next(p for p in (os.path.join(ROOT, "profiles", t + "_kernels.json") for t in ("r1c", "r1b"))
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


def flops_per_step(c, B=1):
    """F_blk = 28 N D^2 + 4 S D^2 + 4 N^2 D + 4 N S D; + patchify, caption projection."""
    D = c["heads"] * c["head_dim"]
    N, S = c["F"] * c["H"] * c["W"], c["S"]
    blk = 28 * N * D * D + 4 * S * D * D + 4 * N * N * D + 4 * N * S * D
    return B * (c["layers"] * blk + 4 * N * 128 * D + 2 * S * (c["caption"] * D + D * D))


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
    """Time one reference DiT block (oracle) at this config's N, S on all host cores, fp32."""
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


def cpu_baseline(c, repeats=1):
    t = cpu_block_seconds(c, repeats)
    layers = c["layers"]
    return dict(value=1.0 / (t * layers), unit="steps/s", cores=os.cpu_count(), kind="port",
                sample=f"1 of {layers} DiT blocks (oracle, torch-CPU fp32, {os.cpu_count()} threads) at N="
                       f"{c['F'] * c['H'] * c['W']}, S={c['S']}: {t:.2f} s/block, extrapolated x{layers}; "
                       f"mlx (the reference's backend) is not installable here, so this is the restated oracle, not MLX",
                seconds_per_block=t)


def run_reference(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times = []
    for i in range(args.warmup + args.steps):
        t = cpu_block_seconds(c, 1)
        if i >= args.warmup:
            times.append(t)
    t = sum(times) / len(times)
    v = 1.0 / (t * c["layers"])
    cb = dict(value=v, unit="steps/s", cores=os.cpu_count(), kind="port",
              sample=f"each step = 1 of {c['layers']} blocks of the oracle at the full N, extrapolated x{c['layers']}")
    print(json.dumps({
        "impl": "reference", "metric": "denoising steps/sec", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["name"]}, "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mlx_available": mlx_available(),
    }))
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU arm
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


def bench_vae(args, dev, rank, world=1):
    """decode_latent of a 9x16x24 latent (65 frames @ 512x768) exactly as the reference schedules it
    (7-frame chunks, overlap 2, cross-fade, uint8), device-timed; plus the conv kernel's roofline.
    world > 1: the chunks are decoded round-robin over the ranks and collected on rank 0 (strong scaling; this
    latent has only two chunks, so at most two ranks have work)."""
    import ctypes as C
    import torch
    from ltx2_b200 import _lib, synthetic
    from ltx2_b200.video_vae import SimpleVideoDecoder, chunk_plan, decode_latent
    vcfg = synthetic.VaeConfig()
    dec = SimpleVideoDecoder(device=dev)
    dec.load_weights(synthetic.iter_vae_weights(vcfg, seed=0, device=dev, dtype=torch.bfloat16))
    assert not dec.missing_weights()
    import torch.distributed as dist
    lat = synthetic.latents((1, 128, 9, 16, 24), seed=43).to(dev)
    frames = 65
    kw = dict(group=dist.group.WORLD, dst=0) if world > 1 else {}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        out = decode_latent(lat, dec, **kw)
    n = max(3, min(args.steps, 8))
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
    assert rank != 0 or out.shape == (frames, 512, 768, 3)
    # e2e: host latent in, uint8 frames on the host out
    lat_h = lat.cpu().pin_memory()
    out_h = torch.empty(frames, 512, 768, 3, dtype=torch.uint8).pin_memory()
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
        t = torch.tensor([ms, ms_e2e, float(launches)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        t2 = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        launches = int(t2[0])
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
    return {
        "metric": "VAE decode frames/sec", "value": frames * 1000.0 / ms, "unit": "frames/s", "ms_per_decode": ms,
        "config": {"workload": "decode_latent, latent 1x128x9x16x24 -> 65 frames @ 512x768, V2.0 decoder stack "
                               "(base 128, 5 res blocks/group), reference chunking 7/2 -> chunks " + str(plan),
                   "parallelism": "single GPU" if world == 1 else
                                  f"chunks round-robin over {world} ranks, collected on rank 0 (strong scaling)",
                   "noise": "decode_noise_scale 0.025 (reference default), timestep 0.05"},
        "gpu_launches": int(launches),
        "e2e": {"value": frames * 1000.0 / ms_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(lat_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel())},
        "roofline": {"bound": "tensor", "kernel": "conv3d_kernel (tcgen05 implicit GEMM, all convs of one 7-frame chunk)",
                     "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"],
                     "traffic": profiled_traffic("prof_conv", "conv3d_kernel<128>"),
                     "traffic_note": "DRAM bytes of one last-stage conv launch (128->128 ch, 49x128x192) from "
                                     "profiles/r1c_kernels.json; algorithmic bytes 330 MB in + 308 MB out + 0.9 MB weights",
                     "launches": int(pl.value), "ms_in_decode": pm.value, "flops_in_decode": pf.value,
                     "decode": {"algorithmic_flops": alg, "achieved": alg / (ms * 1e-3) / 1e12,
                                "frac": alg / (ms * 1e-3) / 1e12 / pk["tf"]}},
    }


def run_ours(args, c):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ltx2_b200 import _lib, synthetic
    from ltx2_b200.transformer import LTXModel, LTXModelType, Modality, X0Model

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    D = c["heads"] * c["head_dim"]
    cfg = synthetic.DitConfig(num_attention_heads=c["heads"], attention_head_dim=c["head_dim"], num_layers=c["layers"],
                              cross_attention_dim=D, caption_channels=c["caption"])
    model = LTXModel(model_type=LTXModelType.VideoOnly, num_attention_heads=c["heads"],
                     attention_head_dim=c["head_dim"], num_layers=c["layers"], cross_attention_dim=D,
                     caption_channels=c["caption"], device=dev)
    from ltx2_b200.loader import iter_engine_weights
    model.load_weights(iter_engine_weights(synthetic.iter_dit_weights(cfg, seed=0, device=dev, dtype=torch.bfloat16),
                                           include_audio=False))
    assert not model.missing_weights()
    x0model = X0Model(model)

    N, S = c["F"] * c["H"] * c["W"], c["S"]
    cp = world > 1 and args.parallel == "cp"
    if cp:
        # ONE sample sharded over the ranks (context parallel): same inputs on every rank
        from ltx2_b200 import context_parallel
        context_parallel.enable(model, batch=1, n_total=N, context_tokens=S)
    lat0 = synthetic.latents((1, N, 128), seed=42 + (0 if cp else rank))
    ctx0 = (synthetic.latents((1, S, c["caption"]), seed=7, std=0.1)).to(torch.bfloat16)
    pos0 = synthetic.video_positions(1, c["F"], c["H"], c["W"], fps=24.0)
    lat_d, ctx_d, pos_d = lat0.to(dev), ctx0.to(dev), pos0.to(dev)
    sig_d = [torch.tensor([s], device=dev) for s in DISTILLED_SIGMAS]
    n_sig = len(DISTILLED_SIGMAS) - 1

    def step_device(i, latent):
        s, s_next = DISTILLED_SIGMAS[i % n_sig], DISTILLED_SIGMAS[i % n_sig + 1]
        x0 = x0model(Modality(latent=latent, context=ctx_d, context_mask=None, timesteps=sig_d[i % n_sig],
                              positions=pos_d))
        # Euler step on x0 (diffusion_steps.py:55-67), host glue
        return latent + (latent - x0) / s * (s_next - s)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    latent = lat_d.clone()
    for i in range(args.warmup):
        latent = step_device(i, latent)
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = _lib.lib().ltx2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    latent = lat_d.clone()
    for i in range(args.steps):
        latent = step_device(i, latent)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib().ltx2_launch_count() - launches0
    clocks = sampler.stop()
    finite = bool(torch.isfinite(latent).all())

    # ---- e2e: host buffers in, host result out, every step ----
    lat_h, ctx_h, pos_h = lat0.pin_memory(), ctx0.pin_memory(), pos0.pin_memory()
    sig_h = [torch.tensor([s]).pin_memory() for s in DISTILLED_SIGMAS]
    out_h = torch.empty(1, N, 128).pin_memory()
    h2d = lat_h.numel() * 4 + ctx_h.numel() * 2 + pos_h.numel() * 4 + 4
    d2h = out_h.numel() * 4

    def step_e2e(i):
        x0 = x0model(Modality(latent=lat_h, context=ctx_h, context_mask=None, timesteps=sig_h[i % n_sig],
                              positions=pos_h))
        out_h.copy_(x0, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the result (mx.eval analogue)

    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)

    # ---- roofline of the dominant kernel class, from one profiled step ----
    L = _lib.lib()
    _lib.check(L.ltx2_dit_set_profile(model._h, 1))
    step_device(0, lat_d.clone())
    pm, pf, pl = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_int64 * 2)()
    _lib.check(L.ltx2_dit_profile_read(model._h, pm, pf, pl, 2))
    _lib.check(L.ltx2_dit_set_profile(model._h, 0))

    # ---- second half of the metric: VAE decode frames/s (65 frames @ 512x768, BASELINE.json configs[4]) ----
    vae = None
    if args.config == "19b" and not args.no_vae:
        if world == 1:
            del x0model, model
            torch.cuda.empty_cache()
        vae = bench_vae(args, dev, rank, world)
    if world > 1:
        dist.barrier()

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    ms_step = ms / args.steps
    samples = 1 if cp else world
    value = samples * 1000.0 / ms_step
    gemm_tf = pf[0] / (pm[0] * 1e-3) / 1e12 if pm[0] > 0 else 0.0
    attn_tf = pf[1] / (pm[1] * 1e-3) / 1e12 if pm[1] > 0 else 0.0
    fl = flops_per_step(c)
    out = {
        "metric": "denoising steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if cp else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": c["name"],
                   "parallelism": (f"cp{world}: token axis sharded, heads re-sharded for self-attention by peer-memory "
                                   f"stores fused into the q/k-norm and attention kernels" if cp else
                                   (f"replicas x{world}" if world > 1 else "single")),
                   "l2": "weights read per step (25.8 GB bf16) exceed the 126 MB L2; no flush needed",
                   "weights": "seeded random init, reference key names", "residual_stream": "fp32",
                   "gemm_operands": "bf16, fp32 accumulate"},
        "clocks": clocks, "gpu_launches": int(launches), "finite": finite,
        "e2e": {"value": samples * 1000.0 * args.steps / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all DiT linears of one step)",
                     "achieved": gemm_tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf"],
                     "traffic": profiled_traffic("prof_gemm", "gemm_bf16_kernel", skip=3),
                     "traffic_note": "mean DRAM bytes per launch over the 5 per-block GEMMs of one block (QKV, attn out, "
                                     "text q, text kv, text out) from profiles/r1c_kernels.json; algorithmic bytes of "
                                     "those launches average 128 MB", "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                     "launches": int(pl[0]), "ms_in_step": pm[0], "flops_in_step": pf[0],
                     "attention": {"kernel": "attention_pair_kernel (tcgen05, two softmax streams per CTA; head_dim 128)",
                                   "achieved": attn_tf, "frac": attn_tf / pk["tf"], "launches": int(pl[1]),
                                   "ms_in_step": pm[1], "flops_in_step": pf[1],
                                   "traffic": profiled_traffic("prof_attn", "attention_pair_kernel"),
                                   "traffic_note": "mean DRAM bytes per launch over one self- and one text "
                                                   "cross-attention launch (profiles/r1c_kernels.json); algorithmic "
                                                   "bytes: Q, K, V, O of the launch = 113 MB / 50 MB"},
                     "step": {"algorithmic_flops": fl, "achieved": fl / (ms_step * 1e-3) / 1e12,
                              "frac": fl / (ms_step * 1e-3) / 1e12 / (pk["tf"] * world),
                              "note": "whole-step FLOPs over all ranks / step time, against world x the per-GPU peak"}},
    }
    if vae is not None:
        out["vae"] = vae
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(c)
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-vae", action="store_true", help="skip the VAE decode leg")
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
