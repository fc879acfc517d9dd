"""CPU restatement (torch fp32) of the elementwise tail of the reference's denoising loops.  TEST INFRASTRUCTURE ONLY:
imported by tests/ (and nothing in the product package).  Pinned by tests/golden/sampling.npz, which was produced by the
reference's own ``EulerDiffusionStep.step``, ``CFGGuider.guide`` and ``post_process_latent`` (tests/golden/make_golden.py).
"""
import torch


def to_velocity(sample, sigma, denoised):
    """core_utils.py:34-62"""
    if float(sigma) == 0:
        raise ValueError("Sigma can't be 0.0")
    return (sample.float() - denoised.float()) / float(sigma)


def euler_step(sample, denoised, sigmas, step_index):
    """components/diffusion_steps.py:36-67"""
    sigma, sigma_next = float(sigmas[step_index]), float(sigmas[step_index + 1])
    return sample.float() + to_velocity(sample, sigma, denoised) * (sigma_next - sigma)


def cfg_guide(cond, uncond, scale):
    """components/guiders.py:40-44"""
    return cond + (scale - 1) * (cond - uncond)


def post_process_latent(denoised, mask, clean):
    """pipelines/common.py:169-190"""
    if mask.ndim == 2 and denoised.ndim == 3:
        mask = mask.unsqueeze(-1)
    return denoised * mask + clean * (1 - mask)
