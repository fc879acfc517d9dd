"""The elementwise tail of the reference's denoising loops on the device, in one fused pass.

Mirrors, with the same names and argument meaning:
  * ``EulerDiffusionStep.step``   components/diffusion_steps.py:36-67  (``to_velocity``: core_utils.py:34-62)
  * ``CFGGuider``                 components/guiders.py:26-47
  * ``post_process_latent``       pipelines/common.py:169-190
and adds ``denoise_update`` (guide -> masked blend -> Euler step in ONE kernel, ``ltx2_denoise_update``) plus
``euler_denoising_loop``, the distilled pipeline's loop (pipelines/distilled.py:214-253) with every tensor resident on
the GPU: per step one X0Model call and one update kernel, and no host synchronisation inside the loop.  Pipelines that keep their own loop can
swap in the three mirrors one by one; all arithmetic is fp32 like the reference's.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from ._lib import check, lib, ptr, stream_ptr
from .transformer import Modality, to_device


def _f32(x, dev) -> torch.Tensor:
    return to_device(x, dev).to(torch.float32).contiguous()


def denoise_update(sample, cond_x0, sigma: float, sigma_next: float, *, uncond_x0=None, cfg_scale: float = 1.0,
                   denoise_mask=None, clean_latent=None, return_denoised: bool = False):
    """sample, cond_x0[, uncond_x0, clean_latent] (B,T,C); denoise_mask (B,T) or (B,T,1) -> next sample (B,T,C) fp32.

    ``sigma == 0`` raises ValueError("Sigma can't be 0.0") like ``to_velocity`` (core_utils.py:54-55)."""
    if float(sigma) == 0.0:
        raise ValueError("Sigma can't be 0.0")
    dev = torch.device("cuda", torch.cuda.current_device())
    x, c = _f32(sample, dev), _f32(cond_x0, dev)
    assert x.shape == c.shape and x.ndim == 3, f"Shape mismatch: {tuple(x.shape)} vs {tuple(c.shape)}"
    u = _f32(uncond_x0, dev) if uncond_x0 is not None else None
    m = cl = None
    if denoise_mask is not None:
        if clean_latent is None:
            raise ValueError("denoise_mask needs clean_latent")
        m = _f32(denoise_mask, dev).reshape(x.shape[0], x.shape[1])
        cl = _f32(clean_latent, dev)
        assert cl.shape == x.shape
    out = torch.empty_like(x)
    den = torch.empty_like(x) if return_denoised else None
    B, T, Cc = x.shape
    check(lib().ltx2_denoise_update(ptr(x), ptr(c), ptr(u), C.c_float(cfg_scale), ptr(m), ptr(cl), C.c_float(sigma),
                                    C.c_float(sigma_next), ptr(out), ptr(den), B * T, Cc, stream_ptr()),
          "ltx2_denoise_update")
    return (out, den) if return_denoised else out


class EulerDiffusionStep:
    """components/diffusion_steps.py:22-67: sample + (sample - denoised) / sigma * (sigma_next - sigma), in fp32."""

    def step(self, sample, denoised_sample, sigmas: Sequence[float], step_index: int) -> torch.Tensor:
        sigma, sigma_next = float(sigmas[step_index]), float(sigmas[step_index + 1])
        return denoise_update(sample, denoised_sample, sigma, sigma_next)


@dataclass(frozen=True)
class CFGGuider:
    """components/guiders.py:26-47."""
    scale: float

    def delta(self, cond, uncond):
        return (self.scale - 1) * (cond - uncond)

    def guide(self, cond, uncond):
        return cond + self.delta(cond, uncond)

    def enabled(self) -> bool:
        return self.scale != 1.0


def post_process_latent(denoised, denoise_mask, clean_latent):
    """pipelines/common.py:169-190 (torch tensors on any device)."""
    if denoise_mask.ndim == 2 and denoised.ndim == 3:
        denoise_mask = denoise_mask.unsqueeze(-1)
    return (denoised * denoise_mask + clean_latent * (1 - denoise_mask)).to(denoised.dtype)


def timestep_classes_from_mask(mask: torch.Tensor):
    """(B, T) denoise mask -> (unit class values (n_cls,), row_cls (B, T) int32): the distinct (batch, mask value) pairs.
    timesteps_from_mask (pipelines/common.py:193-203) is mask * sigma, so for a fixed mask the class structure is the
    same at every step and only the values scale with sigma.  One host synchronisation per LOOP (torch.unique), none per
    step."""
    B, T = mask.shape
    keyed = mask.double() + 4.0 * torch.arange(B, device=mask.device, dtype=torch.float64)[:, None]   # mask in [0, 1]
    uniq, inv = torch.unique(keyed.reshape(-1), return_inverse=True)
    vals = (uniq - 4.0 * torch.floor(uniq / 4.0)).to(torch.float32)
    if vals.numel() > 64:
        return None
    return vals.contiguous(), inv.reshape(B, T).to(torch.int32).contiguous()


def euler_denoising_loop(x0_model, latent, context, positions, sigmas: Sequence[float], *, denoise_mask=None,
                         clean_latent=None, negative_context=None, cfg_scale: float = 1.0,
                         timestep_classes=None) -> torch.Tensor:
    """pipelines/distilled.py:214-253 (video only) with optional CFG (pipelines/one_stage.py:267-326): for every sigma,
    x0 = model(Modality(latent, timesteps = mask * sigma, ...)), then the fused update.  Returns the final latent.

    Nothing in the loop synchronises with the host: without a mask the timesteps are the scalar (B,) form; with a mask
    the (batch, sigma) classes are built once (timestep_classes_from_mask) and only rescaled per step; CFG runs cond and
    uncond as ONE batch-of-2 forward (BASELINE.json north_star), which also keeps the context tensor -- and so the
    engine's cached text K/V -- the same object for every step."""
    dev = torch.device("cuda", torch.cuda.current_device())
    x = _f32(latent, dev)
    B, T, _ = x.shape
    use_mask = denoise_mask is not None
    mask = _f32(denoise_mask, dev).reshape(B, T) if use_mask else None
    clean = _f32(clean_latent, dev) if clean_latent is not None else None
    if use_mask and clean is None:
        raise ValueError("denoise_mask needs clean_latent")
    ctx, pos = to_device(context, dev), to_device(positions, dev)
    cfg = negative_context is not None and cfg_scale != 1.0
    if cfg:
        ctx = torch.cat([ctx, to_device(negative_context, dev)], dim=0).contiguous()      # (2B, S, C): cond | uncond
        pos = torch.cat([pos, pos], dim=0).contiguous()
    rep = 2 if cfg else 1
    classes = None
    if use_mask:      # pre-computed by a caller that must not synchronise here (GraphedDenoiser captures this loop)
        classes = timestep_classes if timestep_classes is not None else timestep_classes_from_mask(mask.repeat(rep, 1))
    for i in range(len(sigmas) - 1):
        sigma = float(sigmas[i])
        sig = torch.full((B * rep,), sigma, device=dev)
        xin = torch.cat([x, x], dim=0) if cfg else x
        if not use_mask:
            mod = Modality(latent=xin, context=ctx, context_mask=None, timesteps=sig, positions=pos, sigma=sig)
        elif classes is not None:
            mod = Modality(latent=xin, context=ctx, context_mask=None, timesteps=sig, positions=pos, sigma=sig,
                           timestep_classes=(classes[0] * sigma, classes[1]))
        else:                                                       # > 64 distinct mask values: per-token form
            mod = Modality(latent=xin, context=ctx, context_mask=None, timesteps=mask.repeat(rep, 1) * sigma,
                           positions=pos, sigma=sig)
        out = x0_model(mod)
        cond, uncond = (out[:B], out[B:]) if cfg else (out, None)
        x = denoise_update(x, cond, sigma, float(sigmas[i + 1]), uncond_x0=uncond, cfg_scale=cfg_scale,
                           denoise_mask=mask, clean_latent=clean)
    return x


class GraphedDenoiser:
    """The whole denoising loop of one sample as ONE CUDA graph ("host denoise loop as a
    CUDA-graph-captured native loop").

    The first call runs the loop once eagerly (allocates the engine workspace, fills caches), captures it -- every X0Model
    forward and every fused update of all steps, about 740 kernel launches per step -- and later calls only copy the new
    latent / context into the static input buffers and replay the graph: no Python, ctypes or launch work per step.  Valid
    because nothing in the loop synchronises with the host (scalar or class timesteps, device-side sigma, the V1 text K/V
    reuse decided once per sample and therefore identical in every replay).  Single-GPU engines only: the cross-GPU barriers
    of the context-parallel forward carry a host-incremented epoch.  Results are bit-identical to the eager loop."""

    def __init__(self, x0_model, sigmas: Sequence[float], *, cfg_scale: float = 1.0):
        self.x0_model, self.sigmas, self.cfg_scale = x0_model, [float(s) for s in sigmas], float(cfg_scale)
        self._graph = None
        self._key = None

    def _loop(self):
        return euler_denoising_loop(self.x0_model, self._lat, self._ctx, self._pos, self.sigmas, denoise_mask=self._mask,
                                    clean_latent=self._clean, negative_context=self._nctx, cfg_scale=self.cfg_scale,
                                    timestep_classes=self._classes)

    def __call__(self, latent, context, positions, *, denoise_mask=None, clean_latent=None, negative_context=None):
        dev = torch.device("cuda", torch.cuda.current_device())
        model = self.x0_model.velocity_model
        if getattr(model, "_cp", None) is not None:
            raise NotImplementedError("GraphedDenoiser: context-parallel engines are not graph-captured")
        ins = dict(lat=_f32(latent, dev), ctx=to_device(context, dev), pos=to_device(positions, dev),
                   mask=_f32(denoise_mask, dev) if denoise_mask is not None else None,
                   clean=_f32(clean_latent, dev) if clean_latent is not None else None,
                   nctx=to_device(negative_context, dev) if negative_context is not None else None)
        # the (batch, mask value) classes are found OUTSIDE the graph (torch.unique synchronises); the graph sees them as
        # static buffers, and a different class count means a different graph
        classes = None
        if ins["mask"] is not None:
            rep = 2 if (ins["nctx"] is not None and self.cfg_scale != 1.0) else 1
            B, T, _ = ins["lat"].shape
            classes = timestep_classes_from_mask(ins["mask"].reshape(B, T).repeat(rep, 1))
            if classes is None:
                raise NotImplementedError("GraphedDenoiser: more than 64 distinct mask values")
        key = tuple((k, None if v is None else (tuple(v.shape), v.dtype)) for k, v in ins.items()) + \
            (None if classes is None else int(classes[0].numel()),)
        if self._graph is not None and key == self._key and classes is not None:
            self._classes[0].copy_(classes[0])
            self._classes[1].copy_(classes[1])
        if self._graph is None or key != self._key:
            self._key = key
            self._classes = None if classes is None else (classes[0].clone(), classes[1].clone())
            for k, v in ins.items():
                setattr(self, "_" + k, None if v is None else v.clone())
            model.reset_context_cache()
            self._loop()                                   # eager warm-up: workspace, tensor maps, caches
            torch.cuda.synchronize()
            model.reset_context_cache()                    # the captured sample starts like a fresh one: miss, then hits
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._out = self._loop()
        else:
            for k, v in ins.items():
                if v is not None:
                    getattr(self, "_" + k).copy_(v)
        self._graph.replay()
        return self._out.clone()
