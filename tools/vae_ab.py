#!/usr/bin/env python
"""A/B of the VAE decode variants on one GPU: fused conv producer (LTX2_VAE_FUSE) x SM-pair conv kernel
(LTX2_CONV_PAIR) x three kernel rows per pipeline stage (LTX2_CONV_KH3), same decoder, same latent.  Prints frames/s of decode_latent (65 frames @ 512x768), the conv class
time of one profiled 7-frame chunk, and the relative difference of each variant's video to the unfused 1-CTA one."""
import ctypes as C
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ltx2_b200 import _lib, synthetic  # noqa: E402
from ltx2_b200.video_vae import SimpleVideoDecoder, decode_latent  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    dec = SimpleVideoDecoder(device=dev)
    dec.load_weights(synthetic.iter_vae_weights(synthetic.VaeConfig(), seed=0, device=dev, dtype=torch.bfloat16))
    lat = synthetic.latents((1, 128, 9, 16, 24), seed=43).to(dev)
    L = _lib.lib()
    base = None
    combos = [("0", "0", "0"), ("0", "2", "0"), ("1", "0", "0"), ("1", "2", "0"), ("1", "2", "1"), ("0", "2", "1"),
              ("1", "2", "0"), ("1", "2", "1")]          # (the last two repeat: clocks drift over the run)
    for fuse, pair, kh3 in combos:
        os.environ["LTX2_VAE_FUSE"], os.environ["LTX2_CONV_PAIR"], os.environ["LTX2_CONV_KH3"] = fuse, pair, kh3
        dec.decode_noise_scale = 0.0
        vid = dec(lat[:, :, :3].contiguous(), timestep=0.05)
        if base is None:
            base = vid.clone()
        diff = float((vid - base).norm() / base.norm())
        dec.decode_noise_scale = 0.025
        for _ in range(3):
            decode_latent(lat, dec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n = 6
        for _ in range(n):
            decode_latent(lat, dec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        _lib.check(L.ltx2_vae_set_profile(dec._h, 1))
        dec(lat[:, :, :7].contiguous(), timestep=0.05)
        pm, pf, pl = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(L.ltx2_vae_profile_read(dec._h, C.byref(pm), C.byref(pf), C.byref(pl)))
        per = []
        for i in range(pl.value):
            a, b = C.c_double(), C.c_double()
            _lib.check(L.ltx2_vae_profile_launch(dec._h, i, C.byref(a), C.byref(b)))
            per.append((a.value, b.value))
        _lib.check(L.ltx2_vae_set_profile(dec._h, 0))
        if os.environ.get("VAE_AB_DETAIL"):
            print("   per launch ms (TF/s): " + " ".join(f"{a * 1e3:.0f}us({b / a / 1e9:.0f})" for a, b in per))
        print(f"fuse={fuse} pair={pair} kh3={kh3}: {65e3 / ms:7.1f} frames/s ({ms:.2f} ms)  conv class {pm.value:.2f} ms "
              f"({pf.value / pm.value / 1e9:.0f} TF/s, {pl.value} launches)  rel diff to unfused/1-CTA {diff:.2e}", flush=True)


if __name__ == "__main__":
    main()
