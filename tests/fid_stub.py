"""Deterministic stand-ins for the Inception network and the image batches of the FID tests (the Inception forward is out
of scope): shared by tests/golden/make_golden.py (which drives the UNMODIFIED reference class with them) and
by the CPU / GPU tests, so that both sides see bit-identical features."""
import math

import torch


class FeatureNet(torch.nn.Module):
    """[B, 3, 16, 16] images -> [B, feature_size] features: fixed random projection + tanh, with a non-zero mean."""

    def __init__(self, feature_size: int, seed: int = 7):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("w", torch.randn(768, feature_size, generator=g) / math.sqrt(768.0))
        self.register_buffer("b", torch.randn(feature_size, generator=g) * 0.3)

    def forward(self, img):
        x = img.float().flatten(1)
        return torch.tanh(3.0 * (x - 0.5) @ self.w + self.b) + 0.25


def images(seed: int, n: int, gain: float = 1.0, offset: float = 0.0) -> torch.Tensor:
    """n synthetic 3x16x16 images in [0, 1] with spatially correlated content (CPU generator: reproducible)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n, 3, 4, 4, generator=g)
    fine = torch.rand(n, 3, 16, 16, generator=g)
    img = 0.6 * torch.nn.functional.interpolate(base, scale_factor=4, mode="nearest") + 0.4 * fine
    return (gain * img + offset).clamp(0.0, 1.0)


CASES = {
    # name: (feature_size, n_generated, n_samples, batch)
    "f64": (64, 1100, 1200, 275),
    "f2048_deficient": (2048, 1200, 1100, 300),      # fewer observations than features: singular covariances
    "f2048_full": (2048, 2304, 2176, 512),
}
