"""Seeded synthetic latents for the parity tests and the benchmark.  Pure torch; no reference data."""
from __future__ import annotations

import math
from typing import Tuple

import torch
from torch import Tensor


def gaussian_spec(d: int, seed: int, kappa: float = 1e2, device="cpu") -> Tuple[Tensor, Tensor]:
    """mu ~ N(0,1)^d and a factor H with Sigma = H H^T = Q diag(lambda) Q^T, lambda log-spaced in [1/kappa, 1]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    mu = torch.randn(d, generator=g, dtype=torch.double)
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=g, dtype=torch.double))
    lam = torch.logspace(-math.log10(kappa), 0.0, d, dtype=torch.double)
    return mu.to(device), (q * lam.sqrt()).to(device)


def gaussian_latents(n: int, d: int, seed: int, kappa: float = 1e2, device="cpu", chunk: int = 1 << 16,
                     shift: float = 0.0, scale: float = 1.0, sample_seed: int = None) -> Tensor:
    """x = mu + H z, fp32 [n, d], generated chunk-wise on `device`.  `seed` fixes the distribution (mu, H);
    `sample_seed` (default: derived from `seed`) the draws - ranks of a data-parallel run share `seed` only."""
    mu, half = gaussian_spec(d, seed, kappa, device)
    mu32, half32 = (mu + shift).float(), (half * scale).float()
    g = torch.Generator(device=device).manual_seed(seed + 7919 if sample_seed is None else sample_seed)
    out = torch.empty(n, d, dtype=torch.float32, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        z = torch.randn(hi - lo, d, generator=g, dtype=torch.float32, device=device)
        out[lo:hi] = z @ half32.T + mu32
    return out


def mixture_latents(n: int, d: int, seed: int, components: int = 10, kappa: float = 1e2, device="cpu") -> Tensor:
    """K-component Gaussian mixture (means ~ N(0, 3^2), per-component covariance as above, uniform weights)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    centers = (3.0 * torch.randn(components, d, generator=g)).to(device)
    per = -(-n // components)
    parts = [gaussian_latents(per, d, seed * 131 + 17 * k + 1, kappa, device) + centers[k] for k in range(components)]
    x = torch.cat(parts)[:n]
    perm = torch.randperm(n, generator=torch.Generator(device="cpu").manual_seed(seed + 1)).to(device)
    return x[perm].contiguous()


def point_clouds(n: int, m: int, d: int, seed: int, device="cpu") -> Tuple[Tensor, Tensor]:
    """Sinkhorn workload (cfg3): source = Gaussian latents, target = 10-component mixture, scaled to O(1) coordinates."""
    x = gaussian_latents(n, d, seed, device=device)
    y = mixture_latents(m, d, seed + 1, device=device) / 3.0
    return x, y
