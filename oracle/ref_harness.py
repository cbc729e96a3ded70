"""Drive the UNMODIFIED reference (/root/reference) under the shim set.  TEST INFRASTRUCTURE.

Only usable in the build container (the reference tree does not exist on the GPU box).  Used by
oracle/gen_golden.py (fixtures under tests/golden/) and by tests that pin the oracle
restatement (oracle/creste_oracle.py) against the reference's own code.
"""
import copy
import os

import yaml

from . import ref_shims

CFG_ROOT = os.path.join(ref_shims.REFERENCE_ROOT, "configs")
SSC_YAML = "model/ssc_sam/terrainnet_supcon_sam2dynelev_jointdinopretrain.yaml"
IRL_YAML = "model/traversability/terrainnet_maxentirlcf_msfcn_sam2dynsemelev.yaml"
DISTILL_YAML = "model/distillation/effnet_ds2_dinov2_128.yaml"


def _load(rel):
    with open(os.path.join(CFG_ROOT, rel)) as f:
        return yaml.safe_load(f)


def compose_cfgs(image_size=(512, 960), map_size=None, action_horizon=None):
    """Plain-dict versions of the three model configs, Hydra composition emulated by hand.

    configs/model/traversability/*.yaml:20-22 nests the ssc_sam model yaml under
    `vision_backbone` (`- ssc_sam@vision_backbone: ...`).
    """
    ssc = _load(SSC_YAML)
    ssc["vision_backbone"]["effnet_cfgs"]["image_size"] = list(image_size)
    irl = _load(IRL_YAML)
    irl.pop("defaults", None)
    irl["vision_backbone"] = copy.deepcopy(ssc)
    if map_size is not None:
        irl["map_size"] = list(map_size)
        for l in irl["loss"]:
            l["map_sz"] = list(map_size)
    if action_horizon is not None:
        irl["action_horizon"] = action_horizon
    dis = _load(DISTILL_YAML)
    dis["vision_backbone"]["effnet_cfgs"]["image_size"] = list(image_size)
    return {"ssc": ssc, "irl": irl, "distill": dis}


def ref_modules():
    """Import the reference's hot-path modules (after installing the shims)."""
    ref_shims.install()
    import creste.models.lfd as lfd
    import creste.models.terrainnet as terrainnet
    import creste.models.distillation as distillation
    import creste.models.blocks.vin as vin
    import creste.models.blocks.splat_projection as splat
    import creste.models.blocks.conv as conv
    import creste.utils.loss_utils as loss_utils
    import creste.utils.train_utils as train_utils
    import creste.utils.projection as projection
    import creste.utils.depth_utils as depth_utils
    return dict(lfd=lfd, terrainnet=terrainnet, distillation=distillation, vin=vin, splat=splat,
                conv=conv, loss_utils=loss_utils, train_utils=train_utils,
                projection=projection, depth_utils=depth_utils)


def build_ref_maxentirl(image_size=(512, 960), solve_mdp=False, map_size=None,
                        action_horizon=None, zero_terminal_state=None):
    mods = ref_modules()
    from omegaconf import OmegaConf
    cfgs = compose_cfgs(image_size, map_size, action_horizon)
    irl = cfgs["irl"]
    irl["solve_mdp"] = solve_mdp
    if zero_terminal_state is not None:
        irl["zero_terminal_state"] = zero_terminal_state
    model = mods["lfd"].MaxEntIRL(OmegaConf.create(irl))
    return model, cfgs


def build_ref_vin():
    mods = ref_modules()
    from omegaconf import OmegaConf
    irl = compose_cfgs()["irl"]
    kw = irl["traversability_head"]["net_kwargs"]
    return mods["vin"].VIN(OmegaConf.create(kw["reward_cfg"]), OmegaConf.create(kw["qvalue_cfg"]))
