"""Import the UNMODIFIED reference package (`ot_vae_lightning`) from a directory that holds it - `/root/reference` in the
authoring container (tests/golden/make_golden.py) or the pip-installed copy under `baseline/_ref` (bench.py --impl
reference; installed by `__graft_entry__.build()`, git-ignored, shipped to the GPU box).

The reference's `utils/__init__.py` imports `pytorch_lightning`, and `metrics/fid.py` imports `torchmetrics`; neither is in
the image (no network).  A meta-path finder serves inert stub modules for those third-party names; the DDP helpers become
identities, exactly what Lightning returns when no process group exists (reference `utils/__init__.py:21-34`).  Nothing of
the reference itself is stubbed or patched.
"""
import importlib.abc
import importlib.machinery
import sys
import types
import warnings


class _Anything:
    """Attribute sink: any attribute / call / subclassing works and does nothing."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


def _metric_base():
    """The part of torchmetrics' `Metric` the reference FID class relies on (metrics/fid.py:66-97): an nn.Module whose
    `add_state` registers the default tensor as an attribute that `update` accumulates into."""
    import torch

    class Metric(torch.nn.Module):
        def __init__(self, **kwargs):
            super().__init__()

        def add_state(self, name, default, dist_reduce_fx=None):
            self.register_buffer(name, default.clone())

    return Metric


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("pytorch_lightning", "torch_ema", "jsonargparse", "torchmetrics", "lovely_tensors", "retrying")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        name = module.__name__
        if name == "pytorch_lightning.utilities.distributed":
            module.sync_ddp_if_available = lambda x, *a, **k: x
            module.gather_all_tensors = lambda x, *a, **k: [x]
            module.distributed_available = lambda: False
        if name == "pytorch_lightning.utilities":
            module.rank_zero_warn = warnings.warn
            module.rank_zero_info = print
            module.rank_zero_only = lambda f: f
        if name == "pytorch_lightning.utilities.apply_func":
            module.apply_to_collection = lambda data, dtype, fn, *a, **k: fn(data)
        if name == "retrying":
            module.retry = lambda *a, **k: (lambda f: f)
        if name == "torchmetrics.metric":
            module.Metric = _metric_base()
        if name == "torchmetrics.image.fid":
            import torch

            class NoTrainInceptionV3(torch.nn.Module):        # never instantiated: the fixtures pass their own `net`
                pass

            def _compute_fid(mu1, sigma1, mu2, sigma2):
                """torchmetrics (third-party, unpinned `torchmetrics>=0.9.2`, absent here): the published v1.x body of
                torchmetrics/image/fid.py::_compute_fid, restated - the only non-reference code on the FID fixture path."""
                a = (mu1 - mu2).square().sum(dim=-1)
                b = sigma1.trace() + sigma2.trace()
                c = torch.linalg.eigvals(sigma1 @ sigma2).sqrt().real.sum(dim=-1)
                return a + b - 2 * c

            module.NoTrainInceptionV3 = NoTrainInceptionV3
            module._compute_fid = _compute_fid


def load_reference(root: str):
    """Returns the reference's `ot_vae_lightning` namespace (found under `root`) with `ot.*`, `utils` importable."""
    if "ot_vae_lightning" in sys.modules and getattr(sys.modules["ot_vae_lightning"], "_is_reference", False):
        return sys.modules["ot_vae_lightning"]
    sys.meta_path.insert(0, _StubFinder())
    pkg = types.ModuleType("ot_vae_lightning")
    pkg.__path__ = [root + "/ot_vae_lightning"]  # skip the star-importing __init__
    pkg._is_reference = True
    sys.modules["ot_vae_lightning"] = pkg
    import ot_vae_lightning.utils  # noqa: F401
    import ot_vae_lightning.ot.matrix_utils  # noqa: F401
    import ot_vae_lightning.ot.w2_utils  # noqa: F401
    import ot_vae_lightning.ot.distribution_models.gaussian_model  # noqa: F401
    import ot_vae_lightning.ot.distribution_models.codebook_model  # noqa: F401
    import ot_vae_lightning.ot.transport.gaussian_transport  # noqa: F401
    import ot_vae_lightning.ot.transport.discrete_transport  # noqa: F401
    return pkg
