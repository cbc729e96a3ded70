"""creste_public_b200 -- B200-native (sm_100a) implementation of CREStE's perception->costmap
+ IRL hot path behind the reference's nn.Module surface.  See DESIGN.md / INTEGRATION.md.

    import creste_public_b200 as cb
    model = cb.build_maxentirl(image_size=(512, 960)).cuda().eval()
    out = model((rgbd, p2p))["traversability_preds"]

The compute lives in csrc/libcreste_b200.so (C ABI: include/creste_b200.h).  There is no CPU or
PyTorch-eager fallback: ops raise if the library is missing or tensors are not on a CUDA device.
"""
import os
import sys

from .engine import get_precision, set_precision  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))


def install_as_creste(reference_root=None):
    """Make `import creste.models...` resolve to the sm_100a-backed mirror (drop-in use from the reference's train
    scripts): puts this package directory first on sys.path.

    reference_root: path of a ut-amrl/creste_public checkout.  Given, the mirror is OVERLAID on it: the mirror's
    packages (`creste`, `creste.models`, `creste.models.blocks`, `creste.models.losses`, `creste.utils`,
    `creste.datasets`) get the
    reference's directories appended to their __path__, so every module the mirror provides shadows the
    reference's, and everything else the train scripts import (`creste.datasets.*`, `creste.utils.visualization`,
    `creste.utils.tb_utils`, ...) falls through to the reference tree unchanged.  `<reference_root>/creste` is also
    put on sys.path for the scripts' un-prefixed imports (`from datasets.dataloader import ...`)."""
    import importlib
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    for k in [k for k in sys.modules if k == "creste" or k.startswith("creste.")]:
        del sys.modules[k]
    if reference_root is None:
        return
    ref_pkg = os.path.join(reference_root, "creste")
    if not os.path.isdir(ref_pkg):
        raise FileNotFoundError(ref_pkg)
    from .creste import _overlay
    _overlay.REFERENCE_ROOT = reference_root
    _overlay._loaded.clear()
    for name in ("creste", "creste.models", "creste.models.blocks", "creste.models.losses", "creste.utils",
                 "creste.datasets"):
        pkg = importlib.import_module(name)
        extra = os.path.join(reference_root, *name.split("."))
        if os.path.isdir(extra) and extra not in pkg.__path__:
            pkg.__path__.append(extra)
    # the scripts run with <reference_root>/creste as their script directory, i.e. ahead of site-packages (an
    # installed `datasets` package must not win over the reference's `datasets/`)
    if ref_pkg in sys.path:
        sys.path.remove(ref_pkg)
    sys.path.insert(sys.path.index(_HERE) + 1, ref_pkg)
    # `datasets/` of the reference is a namespace package (no __init__.py): an installed regular package of the same
    # name (HuggingFace datasets) would win whatever the path order, so it is bound explicitly
    import types
    ds = os.path.join(ref_pkg, "datasets")
    if os.path.isdir(ds) and not os.path.isfile(os.path.join(ds, "__init__.py")):
        for k in [k for k in sys.modules if k == "datasets" or k.startswith("datasets.")]:
            del sys.modules[k]
        m = types.ModuleType("datasets")
        m.__path__ = [ds]
        sys.modules["datasets"] = m


def build_maxentirl(cfg=None, image_size=(512, 960), solve_mdp=False, map_size=(64, 128),
                    action_horizon=50):
    """MaxEntIRL (reference creste/models/lfd.py) from a composed config (DictConfig / dict) or
    the shipped defaults."""
    from . import configs
    from .config import as_cfg
    from .creste.models.lfd import MaxEntIRL
    if cfg is None:
        cfg = configs.irl_cfg(image_size, map_size, solve_mdp, action_horizon)
    return MaxEntIRL(as_cfg(cfg))


def build_terrainnet(cfg=None, image_size=(512, 960)):
    from . import configs
    from .config import as_cfg
    from .creste.models.terrainnet import TerrainNet
    return TerrainNet(as_cfg(cfg if cfg is not None else configs.ssc_cfg(image_size)))
