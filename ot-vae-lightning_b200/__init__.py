"""ot_vae_lightning_b200: B200-native latent optimal-transport path of theoad/ot-vae-lightning.

Same Python surface as the reference's `ot_vae_lightning.ot` / `ot_vae_lightning.metrics` for the hot path
(streaming mean/cov statistics, Gaussian W2 machinery, transport map, log-domain Sinkhorn); the arithmetic is
hand-written sm_100a CUDA in `csrc/` behind the C ABI of `include/otk.h`.  No CPU fallback.
"""
__version__ = "0.1.0"


def install_as_reference(name: str = "ot_vae_lightning") -> None:
    """Alias this package's `ot` / `metrics` / `utils` sub-modules under the reference's module paths, so that
    unmodified reference code (`from ot_vae_lightning.ot.w2_utils import ...`) resolves to the B200 kernels.
    See INTEGRATION.md."""
    import importlib
    import sys
    import types
    root = sys.modules.get(name)
    if root is None:
        root = types.ModuleType(name)
        root.__path__ = []
        sys.modules[name] = root
    for sub in ("utils", "ot", "ot.matrix_utils", "ot.w2_utils", "ot.distribution_models", "ot.distribution_models.base",
                "ot.distribution_models.gaussian_model", "ot.distribution_models.codebook_model",
                "ot.distribution_models.gassian_mixture_model", "ot.transport",
                "ot.transport.base", "ot.transport.gaussian_transport", "ot.transport.discrete_transport",
                "ot.transport.gmm_transport",
                "metrics", "metrics.fid"):
        mod = importlib.import_module(f"{__name__}.{sub}")
        sys.modules[f"{name}.{sub}"] = mod
        parent, _, leaf = sub.rpartition(".")
        setattr(sys.modules[f"{name}.{parent}"] if parent else root, leaf, mod)
