#!/usr/bin/env python
"""Multi-GPU VAE decode check, run under torchrun with one rank per GPU: chunk-sharded decode_latent and tile-sharded
decode_tiled must reproduce the single-GPU results (chunks: bit-exact; tiles: fp32 sum order may differ by an ulp)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ltx2_b200 import synthetic  # noqa: E402
from ltx2_b200.tiling import SpatialTilingConfig, TemporalTilingConfig, TilingConfig, decode_tiled  # noqa: E402
from ltx2_b200.video_vae import (SimpleVideoDecoder, decode_latent, decode_latent_video, decode_sharded,  # noqa: E402
                                 disable_temporal_shards, enable_temporal_shards, shard_frames)

BLOCKS = [["res_x", {"num_layers": 2}], ["compress_all", {"multiplier": 2, "residual": True}],
          ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
          ["res_x", {"num_layers": 1}], ["compress_all", {"multiplier": 2, "residual": True}],
          ["res_x", {"num_layers": 1}]]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    G = dist.group.WORLD
    cfg = synthetic.VaeConfig(decoder_blocks=BLOCKS, base_channels=64, timestep_conditioning=True)
    dec = SimpleVideoDecoder(decoder_blocks=BLOCKS, base_channels=64, timestep_conditioning=True, device=dev)
    dec.load_weights(synthetic.vae_weights(cfg, seed=24))
    dec.decode_noise_scale = 0.0
    ok = True
    for T in (9, 16, 23):
        lat = synthetic.latents((1, 128, T, 2, 2), seed=400 + T)
        ref = decode_latent_video(lat, dec)
        out = decode_latent_video(lat, dec, group=G)
        good = torch.equal(out, ref)
        to0 = decode_latent(lat, dec, group=G, dst=0)
        if rank == 0:
            good = good and torch.equal(to0, decode_latent(lat, dec))
        else:
            good = good and to0 is None
        ok = ok and good
        print(f"rank {rank} chunks T_lat={T}: {'OK' if good else 'MISMATCH'}", flush=True)
    lat = synthetic.latents((1, 128, 3, 6, 6), seed=410)
    for name, tc in [("spatial", TilingConfig(SpatialTilingConfig(64, 32), None)),
                     ("spatial+temporal", TilingConfig(SpatialTilingConfig(96, 32), TemporalTilingConfig(16, 8)))]:
        ref = next(decode_tiled(lat, dec, tc, timestep=0.05))
        out = next(decode_tiled(lat, dec, tc, timestep=0.05, group=G))
        d = float((out - ref).abs().max())
        good = out.shape == ref.shape and d <= 1e-5
        ok = ok and good
        print(f"rank {rank} tiles {name}: max|diff| {d:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
    # ---- temporal shards: ONE clip's frames split over the ranks, halo frames exchanged through peer memory ----
    # must be BIT-IDENTICAL to the single-GPU decode (same kernels per frame; only who computes which frame changes)
    single = {}
    cases = [(1, 7, 2, 3), (2, 5, 3, 2), (1, 1, 2, 2), (1, 2, 4, 4), (1, 16, 2, 2)]
    for (B, T, H, W) in cases:
        lat = synthetic.latents((B, 128, T, H, W), seed=500 + T)
        single[(B, T, H, W)] = (lat, dec(lat, timestep=0.05))
    long_lat = synthetic.latents((1, 128, 16, 2, 3), seed=520)
    long_ref_u8 = decode_latent(long_lat, dec)
    long_ref = decode_latent_video(long_lat, dec)
    enable_temporal_shards(dec, (2, 128, 16, 4, 4), group=G)
    for key, (lat, ref) in single.items():
        for dst in (None, 0):
            out = decode_sharded(dec, lat, 0.05, dst=dst)
            good = (out is None) if (dst is not None and rank != dst) else torch.equal(out, ref)
            ok = ok and good
            spans = [shard_frames(dec, key[1], r, world) for r in range(world)]
            print(f"rank {rank} temporal shards B,T,H,W={key} dst={dst} spans={spans}: {'OK' if good else 'MISMATCH'}",
                  flush=True)
    # rank groups: each half of the ranks decodes its own clip at the same time, both collected on rank 0 / everywhere
    if world >= 2 and world % 2 == 0:
        from ltx2_b200.video_vae import collect_clip
        hw = world // 2
        keys = [(1, 7, 2, 3), (2, 5, 3, 2)]
        for dst in (None, 0):
            for rnd in range(3):                      # three rounds: the clip slots alternate and are reused
                s0 = 2 * (rnd % 2)
                for j, key in enumerate(keys):
                    decode_sharded(dec, single[key][0], 0.05, dst=dst, group_ranks=(j * hw, hw), slot=s0 + j, collect=False)
                good = True
                for j, key in enumerate(keys):
                    out = collect_clip(dec, single[key][0].shape, dst, s0 + j)
                    good = good and ((out is None) if (dst is not None and rank != dst) else torch.equal(out, single[key][1]))
                ok = ok and good
                print(f"rank {rank} two rank groups side by side, round {rnd}, dst={dst}: {'OK' if good else 'MISMATCH'}",
                      flush=True)
    # the reference's chunked decode_latent on top of the shards (every chunk split over all ranks), uint8 frames
    out = decode_latent_video(long_lat, dec, group=G)
    good = torch.equal(out, long_ref)
    u8 = decode_latent(long_lat, dec, group=G, dst=0)
    good = good and ((u8 is None) if rank != 0 else torch.equal(u8, long_ref_u8))
    ok = ok and good
    print(f"rank {rank} decode_latent over temporal shards: {'OK' if good else 'MISMATCH'}", flush=True)
    # noise injection: the same noise on every rank (broadcast from rank 0) -> shards agree with each other
    dec.decode_noise_scale = 0.025
    a = decode_sharded(dec, single[(1, 7, 2, 3)][0], 0.05, dst=None)
    parts = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(parts, a)
    good = all(torch.equal(p, parts[0]) for p in parts) and bool(torch.isfinite(a).all())
    ok = ok and good
    print(f"rank {rank} temporal shards with noise injection, same clip on all ranks: {'OK' if good else 'MISMATCH'}", flush=True)
    dec.decode_noise_scale = 0.0
    disable_temporal_shards(dec)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    print("VAE_CP_CHECK_PASS" if int(t) == 1 else "VAE_CP_CHECK_FAIL", flush=True)
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
