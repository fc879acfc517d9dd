"""Context parallelism for the DiT over the GPUs of one NVLink/NVSwitch box.

The reference is single-device, so there is no reference API to mirror; the design keeps the pipelines unchanged:
every rank runs the same sampling loop on the same (replicated) inputs and calls ``transformer(video_modality)``
with the FULL token sequence.  Inside the call

  1. the modality is sliced to this rank's contiguous token range (patchify order is (f,h,w) row-major,
     components/patchifiers.py:94-100, so a slice is a band of latent frames/rows),
  2. the engine runs with tokens sharded; inside self-attention heads are re-sharded by stores to peer memory that
     are fused into the producing kernels (csrc/rowops.cu: qkv_head_scatter, csrc/attention_sm100.cu epilogue),
  3. the (B, N/P, 128) outputs are all-gathered once per call (torch.distributed, NCCL on GPU) so every rank
     returns the full velocity / denoised sample, exactly like the single-GPU model.

Only host logic lives here (slicing, handle exchange, the final gather); it is exercised on CPU with the gloo backend
in tests/test_cp_cpu.py.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def token_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    if n_total % world != 0:
        raise ValueError(f"{n_total} tokens do not split evenly over {world} ranks")
    n = n_total // world
    return rank * n, (rank + 1) * n


def slice_tokens(latent: torch.Tensor, timesteps: torch.Tensor, positions: torch.Tensor, rank: int, world: int):
    """latent (B,N,C), timesteps (B,1)|(B,N), positions (B,n_dims,N,2) -> this rank's token slice (contiguous)."""
    a, b = token_range(latent.shape[1], rank, world)
    ts = timesteps if timesteps.shape[1] == 1 else timesteps[:, a:b]
    return latent[:, a:b].contiguous(), ts.contiguous(), positions[:, :, a:b].contiguous()


def gather_tokens(local: torch.Tensor, group=None) -> torch.Tensor:
    """(B, N/P, C) on every rank -> (B, N, C) on every rank, rank-major token order."""
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=1)


def exchange_handles(handle: bytes, group=None) -> bytes:
    """All-gather the 64-byte CUDA IPC handles (host side, any backend)."""
    world = dist.get_world_size(group)
    out: List[Optional[bytes]] = [None] * world
    dist.all_gather_object(out, handle, group=group)
    assert all(isinstance(h, (bytes, bytearray)) and len(h) == 64 for h in out)
    return b"".join(out)


def enable(model, batch: int, n_total: int, context_tokens: int = 0, group=None) -> None:
    """Turn on context parallelism for an ltx2_b200 LTXModel across `group` (default: the world)."""
    from ._lib import check, lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    buf = C.create_string_buffer(64)
    with torch.cuda.device(model.device):
        check(lib().ltx2_dit_cp_init(model._h, rank, world, batch, n_total, context_tokens, buf), "ltx2_dit_cp_init")
        handles = exchange_handles(buf.raw, group)
        check(lib().ltx2_dit_cp_connect(model._h, handles), "ltx2_dit_cp_connect")
        torch.cuda.synchronize()
    dist.barrier(group=group)
    model._cp = (rank, world, group, batch, n_total)


def set_split_k(model, max_splits: int) -> None:
    """1 = no split-K in the residual GEMMs (sharded forward bit-identical to the single-GPU one), 0 = default."""
    from ._lib import check, lib
    check(lib().ltx2_dit_cp_set_split_k(model._h, int(max_splits)), "ltx2_dit_cp_set_split_k")


def disable(model) -> None:
    """Collective teardown of the exchange region (every rank of the group calls it): close the imported peer mappings,
    barrier, free the own region.  The model is single-GPU afterwards."""
    from ._lib import check, lib
    if model._cp is None:
        return
    group = model._cp[2]
    with torch.cuda.device(model.device):
        check(lib().ltx2_dit_cp_shutdown(model._h, 0), "ltx2_dit_cp_shutdown")
        dist.barrier(group=group)
        check(lib().ltx2_dit_cp_shutdown(model._h, 1), "ltx2_dit_cp_shutdown")
    dist.barrier(group=group)
    model._cp = None
