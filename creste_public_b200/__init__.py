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


def install_as_creste():
    """Make `import creste.models...` resolve to the sm_100a-backed mirror (drop-in use from the
    reference's train scripts): puts this package directory first on sys.path."""
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    for k in [k for k in sys.modules if k == "creste" or k.startswith("creste.")]:
        del sys.modules[k]


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
