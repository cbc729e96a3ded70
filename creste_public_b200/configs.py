"""Default model configs of the shipped reference stages, as plain dicts.

Same structure and values as the reference's Hydra tree (configs/model/ssc_sam/
terrainnet_supcon_sam2dynelev_jointdinopretrain.yaml, configs/model/traversability/
terrainnet_maxentirlcf_msfcn_sam2dynsemelev.yaml with `ssc_sam@vision_backbone` composed in,
configs/model/distillation/effnet_ds2_dinov2_128.yaml), model sections only.  A user of the
reference passes their own composed DictConfig instead; these defaults exist so benchmarks and
tests can build the models where the reference tree (and Hydra) is absent.
"""
import copy

DISCRETIZE = {"mode": "UD", "num_bins": 128, "depth_min": 300, "depth_max": 25600}


def ssc_cfg(image_size=(512, 612)):
    return {
        "project_name": "TerrainNetSAM",
        "run_name": "depth128UD_jointdinopretrain_sam2dynelev_supcon_joint",
        "load_setting": "strict", "use_temporal": False, "use_movability": False,
        "multiview_distillation": False, "depth_embed_dim": 256, "fdn_embed_dim": 128,
        "num_depth_bins": 128, "inpainting_sam_dim": 32, "num_obj_class": 6,
        "weights_path": "", "views": 1,
        "discretize": copy.deepcopy(DISCRETIZE),
        "vision_backbone": {
            "class_name": "DistillationBackbone", "name": "efficientnet-b0", "input_type": "rgbd",
            "weights_path": "", "return_feats": True,
            "effnet_cfgs": {"in_channels": 4, "out_channels": 256, "downsample": 4,
                            "image_size": list(image_size)}},
        "camera_projector": {
            "name": "Cam2MapMulti", "voxel_size": [0.1, 0.1, 3],
            "point_cloud_range": [-12.8, -12.8, -2, 12.8, 12.8, 1], "embed_z": True,
            "z_embed_dim": 32, "z_embed_mode": "mlp", "num_cams": 1,
            "splat_key": "depth_preds_feats",
            "vision_fusion": {"name": "ConvEncoder", "dims": [288, 96], "kernels": [1],
                              "paddings": [0], "norm_type": "batch_norm"}},
        "depth_head": {"name": "depthconv-head", "dims": [256, 128], "kernels": [3], "paddings": [1],
                       "norm_type": "batch_norm"},
        "distillation_head": {"name": "distillation-head", "feature_head": {
            "name": "MultiLayerConv", "kernels": [1, 1, 1], "paddings": [0, 0, 0],
            "dims": [256, 128, 128, 128], "norm_type": "batch_norm"}},
        "bev_classifier": {"name": "InpaintingResNet18MultiHead", "net_kwargs": {
            "input_key": "bev_features", "num_input_features": 96, "num_classes": [32, 6, 2],
            "output_prefix": ["inpainting_sam", "inpainting_sam_dynamic", "elevation"]}},
    }


def ssc_train_cfg(image_size=(512, 612), class_weights=None):
    """ssc_cfg + the training sections of terrainnet_supcon_sam2dynelev_jointdinopretrain.yaml (stage 2,
    train_ssc.py): optimizer / scheduler / the six losses.  `class_weights`: path of the class-frequency text file
    of the dynamic-object CrossEntropy (the reference ships it with its dataset; None = unweighted)."""
    cfg = ssc_cfg(image_size)
    disc = copy.deepcopy(DISCRETIZE)
    ce = {"name": "CrossEntropy", "weight": 2.0, "pred_key": "outputs/inpainting_sam_dynamic_preds",
          "lab_key": "inputs/3d_sam_dynamic_label", "num_class": 6, "class_dim": 1, "task": "joint"}
    if class_weights is not None:
        ce["class_weights"] = class_weights
    cfg.update({
        "batch_size": 8,
        "optimizer": {"name": "Adam", "beta1": 0.9, "beta2": 0.999, "lr": 0.0005},
        "lr_scheduler": {"name": "ExponentialLR", "gamma": 0.98},
        "loss": [
            {"name": "SupPixelConLoss", "views": 1, "weight": 1.0, "pred_key": "outputs/inpainting_sam_preds",
             "lab_key": "inputs/3d_sam_label", "ignore_index": 0, "temperature": 0.1, "task": "joint",
             "contrast_mode": "batch_all"},
            ce,
            {"name": "MSELoss", "weight": 2.0, "pred_key": "outputs/dino_pe_feats", "lab_key": "inputs/fimg_label",
             "overlap_only": False},
            {"name": "CrossEntropyDepth", "weight": 0.5, "pred_key": "outputs/depth_preds_logits",
             "lab_key": "inputs/depth_label", "discretize": copy.deepcopy(disc)},
            {"name": "SmoothL1Depth", "weight": 0.1, "pred_key": "outputs/depth_preds_metric",
             "lab_key": "inputs/depth_label", "beta": 0.5, "discretize": copy.deepcopy(disc)},
            {"name": "SmoothL1", "weight": 3.0, "beta": 0.2, "pred_key": "outputs/elevation_preds",
             "lab_key": "inputs/elevation_label", "absolute": False, "task": "joint"},
        ]})
    return cfg


def irl_cfg(image_size=(512, 612), map_size=(64, 128), solve_mdp=True, action_horizon=50):
    def stack(dims, kernels):
        return {"dims": dims, "kernels": kernels, "stride": [1] * len(kernels),
                "norm_type": "batch_norm"}
    return {
        "project_name": "TraversabilityLearning",
        "run_name": "terrainnet_dinopretrain_maxentirlcf_msfcn_sam2semelev",
        "ckpt_path": "", "weights_path": "", "load_strict": True, "freeze_weights": True,
        "map_ds": 2, "views": 1, "action_horizon": action_horizon, "zero_terminal_state": False,
        "policy_method": "pp", "policy_kwargs": {"method": "sharpen", "temperature": 0.005},
        "solve_mdp": solve_mdp, "map_size": list(map_size),
        "vision_backbone": ssc_cfg(image_size),
        "traversability_head": {
            "name": "MaxEntIRL", "value_iterator": "VIN", "feats_dim": 40, "map_size": 128,
            "policy_method": "pp",
            "net_kwargs": {
                "reward_cfg": {
                    "name": "MultiScaleFCN", "ds": 2,
                    "input_keys": ["inpainting_sam_preds", "inpainting_sam_dynamic_preds",
                                   "elevation_preds"],
                    "output_prefix": ["traversability_preds"],
                    "net_kwargs": {"prepool": stack([40, 64, 32], [5, 3]),
                                   "skip": stack([32, 32, 16], [3, 1]),
                                   "trunk": stack([32, 32, 32], [3, 1]),
                                   "postpool": stack([48, 1], [1])}},
                "qvalue_cfg": {"dims": [1, 8], "kernels": [3], "stride": [1], "padding": [1],
                               "input_keys": ["traversability"], "norm_type": "batch_norm",
                               "discount": 0.99}}},
        "loss": [{"name": "MaxEntIRLLoss", "weight": 1.0, "map_ds": 2, "map_sz": list(map_size),
                  "maxent_weight": 1.0, "reward_weight": 0.01, "alpha": 0.5, "use_fov_mask": True,
                  "pred_key": "outputs/exp_svf", "fov_key": "inputs/fov_mask",
                  "lab_key": "inputs/traversability_label",
                  "cf_key": "inputs/counterfactuals_label"}],
    }


def distill_cfg(image_size=(512, 960)):
    """configs/model/distillation/effnet_ds2_dinov2_128.yaml (stage 1, train_pefree.py): model
    sections + optimizer / scheduler / loss lists."""
    disc = copy.deepcopy(DISCRETIZE)
    base = ssc_cfg(image_size)
    return {
        "project_name": "Distillation", "run_name": "effnet_ds2_dinov2_128",
        "multiview_distillation": False, "weights_path": "", "ckpt_path": "",
        "discretize": disc,
        "vision_backbone": copy.deepcopy(base["vision_backbone"]),
        "depth_head": copy.deepcopy(base["depth_head"]),
        "distillation_head": copy.deepcopy(base["distillation_head"]),
        "batch_size": 4,
        "optimizer": {"name": "Adam", "beta1": 0.9, "beta2": 0.999, "lr": 0.0005, "eps": 1e-7},
        "lr_scheduler": {"name": "ExponentialLR", "gamma": 0.98},
        "loss": [
            {"name": "CrossEntropyDepth", "weight": 0.5, "pred_key": "outputs/depth_preds_logits",
             "lab_key": "inputs/depth_label", "discretize": copy.deepcopy(disc)},
            {"name": "SmoothL1Depth", "weight": 0.1, "pred_key": "outputs/depth_preds_bins",
             "lab_key": "inputs/depth_label", "beta": 0.5, "discretize": copy.deepcopy(disc)},
            {"name": "MSELoss", "weight": 1.0, "pred_key": "outputs/dino_pe_feats",
             "lab_key": "inputs/fimg_label", "overlap_only": False},
        ],
    }
