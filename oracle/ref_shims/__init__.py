"""Import shims that let the UNMODIFIED reference (/root/reference) run in this container.

TEST INFRASTRUCTURE ONLY.  Nothing here ships in the product path; it exists so that
`oracle/gen_golden.py` and `tests/test_oracle_cpu.py` / `tests/test_stage2_cpu.py` / `tests/test_dropin_cpu.py` can execute the reference's
own Python code (creste.models.*, creste.utils.*) to pin the oracle restatement.

The reference imports seven third-party packages that are not installed here (and there is
no network): omegaconf, pytorch_lightning, hydra, kornia, matplotlib, vispy, open3d, shapely,
torch_scatter, efficientnet_pytorch.  We provide:

  * an auto-mock meta-path finder for the visual / IO / Lightning packages (only touched by
    debug and training-wrapper code paths),
  * `omegaconf`  -> oracle/ref_shims/omegaconf_shim.py  (DictConfig / OmegaConf / open_dict),
  * `torch_scatter` -> scatter() over Tensor.scatter_reduce(include_self=False),
  * `efficientnet_pytorch` -> oracle/ref_shims/efficientnet_shim.py, a restatement of the
    public lukemelas/EfficientNet-PyTorch 0.7.1 B0 model (the reference does not vendor or pin
    it: SURVEY.md section 8(c) -- parity at that boundary is UNPINNED by the reference).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("CRESTE_REFERENCE_ROOT", "/root/reference")

_MOCK_ROOTS = {
    "pytorch_lightning", "lightning", "hydra", "kornia", "matplotlib", "vispy", "open3d",
    "shapely", "wandb", "cuml", "flask", "timm", "tensorboard", "seaborn",
}


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []          # behave like a package so sub-imports resolve
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        return None


class _MockFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _MOCK_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, _MockLoader(), is_package=True)
        return None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "creste", "models"))


_installed = False


def install_lightning():
    """Replace the auto-mocked pytorch_lightning by a stand-in whose LightningModule is a real nn.Module, so that the
    reference's train scripts can be imported and their step methods driven (tests/test_dropin_cpu.py)."""
    from . import lightning_shim
    _MOCK_ROOTS.discard("pytorch_lightning")
    for k in [k for k in sys.modules if k == "pytorch_lightning" or k.startswith("pytorch_lightning.")]:
        del sys.modules[k]
    return lightning_shim.install()


def install():
    """Idempotently install the shim set and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    here = os.path.dirname(os.path.abspath(__file__))

    def _absent(name):
        try:
            __import__(name)
            return False
        except Exception:
            return True

    for root in sorted(_MOCK_ROOTS):
        if not _absent(root):
            _MOCK_ROOTS.discard(root)
    sys.meta_path.append(_MockFinder())

    if _absent("omegaconf"):
        from . import omegaconf_shim
        sys.modules["omegaconf"] = omegaconf_shim
    if _absent("torch_scatter"):
        from . import torch_scatter_shim
        sys.modules["torch_scatter"] = torch_scatter_shim
    if _absent("efficientnet_pytorch"):
        from . import efficientnet_shim
        sys.modules["efficientnet_pytorch"] = efficientnet_shim
        sys.modules["efficientnet_pytorch.utils"] = efficientnet_shim.utils

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
