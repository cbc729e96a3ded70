"""Mirror of reference creste/models/blocks/splat_projection.py: Camera2World (:12-51) and
Camera2MapMulti (:53-354), eval mode, scatter_mode='mean', mode='bilinear'.

Kernels (include/creste_b200.h): creste_frustum_to_bev (un-projection, bounds mask, voxel
coordinates -- bit-exact indices), creste_zmlp_concat (z-MLP + concat), creste_conv2d (1x1 fusion
conv + BN + ReLU), creste_splat_soft (bilinear scatter-add + mean normalisation).
"""
import torch
from torch import nn

from creste_public_b200 import engine, ops
from creste_public_b200.engine import PackCache, require_eval
from .conv import ConvEncoder


class Camera2World(nn.Module):
    """depth [B,N,H,W] + p2p [B,N,4,4] -> xyz [B,N,3,H,W] in the LiDAR frame."""

    def forward(self, x):
        """(depth [B,N,H,W], p2p [B,N,4,4]) -> xyz [B,N,3,H,W] (reference :19-51).  The fused hot path
        (Camera2MapMulti.forward_nhwc -> creste_frustum_to_bev) never materialises xyz; this is the
        stand-alone module call, one launch of creste_camera_to_world."""
        depth, p2p = x
        B, N, H, W = depth.shape
        xyz = ops.camera_to_world(depth.reshape(B * N, H, W).float(), p2p.reshape(B * N, 4, 4).float())
        return xyz.view(B, N, 3, H, W)


class Camera2MapMulti(nn.Module):
    def __init__(self, model_cfg, mode="bilinear", scatter_mode="mean"):
        super().__init__()
        self.model_cfg = model_cfg
        self.register_buffer("point_cloud_range", torch.tensor(model_cfg.point_cloud_range))
        self.register_buffer("max_bound", self.point_cloud_range[3:].reshape(1, -1))
        self.register_buffer("min_bound", self.point_cloud_range[:3].reshape(1, -1))
        self.register_buffer("voxel_size", torch.tensor(model_cfg.voxel_size))
        self.register_buffer("grid_size", ((self.point_cloud_range[3:] - self.point_cloud_range[:3])
                                           / self.voxel_size).long())
        self.register_buffer("lidar2map", torch.tensor([
            [0, -1, 0, -self.min_bound[0, 0]], [-1, 0, 0, -self.min_bound[0, 1]],
            [0, 0, -1, -self.min_bound[0, 2]], [0, 0, 0, 1]]).float())
        if mode != "bilinear":
            raise Exception("Unknown splat mode:", mode)
        if scatter_mode != "mean":
            raise NotImplementedError("only scatter_mode='mean' is on the hot path (the 'max' branch "
                                      "is used by the disabled multiview distillation only)")
        self.mode, self.scatter_mode, self.min_weight = mode, scatter_mode, 1.0
        self.NC = model_cfg.get("num_cams", 2)
        self.cam2world = Camera2World()
        if model_cfg["z_embed_mode"] != "mlp":
            raise Exception("Unknown z_embed_mode:", model_cfg["z_embed_mode"])
        zd = model_cfg["z_embed_dim"]
        self.z_proj = nn.Sequential(nn.Linear(1, zd * 2, bias=True), nn.ReLU(),
                                    nn.Linear(zd * 2, zd, bias=True), nn.ReLU())
        self.vision_fusion = ConvEncoder(model_cfg.vision_fusion)
        self._cache = PackCache()
        self._host_geom = None

    def _geom(self):
        """point_cloud_range / voxel_size / grid as host floats (read once; they are buffers)."""
        if self._host_geom is None:
            self._host_geom = ([float(v) for v in self.point_cloud_range.tolist()],
                               [float(v) for v in self.voxel_size.tolist()],
                               [int(v) for v in self.grid_size.tolist()])
        return self._host_geom

    def frustum(self, depth, p2p):
        """depth [M,Hs,Ws], p2p [M,4,4] -> xy [M,P,2], z [M,P], mask [M,P]."""
        rng, vox, _ = self._geom()
        return ops.frustum_to_bev(depth, p2p, rng, vox)

    def forward_nhwc(self, depth, feats_nhwc, p2p, want_nchw=True):
        """depth [M,Hs,Ws] (m), feats NHWC [M,Hs,Ws,F], p2p [M,4,4]; NC == 1 (one camera per
        BEV map, `num_cams: 1` in the shipped configs)."""
        if self.NC != 1:
            raise NotImplementedError("num_cams != 1 is not used by the shipped configs")
        if self.training:
            return self.forward_train(depth, feats_nhwc, p2p, want_nchw)
        if self.z_proj[0].out_features != 64 or self.z_proj[2].out_features != 32:
            raise NotImplementedError("creste_zmlp_concat is specialised for z_embed_dim = 32")
        M, Hs, Ws, F = feats_nhwc.shape
        _, _, grid = self._geom()
        xy, z, mask = self.frustum(depth, p2p)
        l0, l2 = self.z_proj[0], self.z_proj[2]
        w1, b1, w2, b2 = self._zmlp_weights(l0, l2)
        # the maxima travel with the tensors (no amax passes in front of the tensor-core convs): the z-MLP kernel
        # publishes max|cat|; the splat is a weighted mean (sum w f / max(sum w, min_weight), min_weight > 0), so the
        # fusion conv's max|f| bounds the BEV map too
        track = engine.get_precision() in ("3xfp16", "fp16") and not torch.jit.is_tracing()
        amax = torch.empty(1, device=feats_nhwc.device) if track else None
        cat = ops.zmlp_concat(feats_nhwc, z, w1, b1, w2, b2, **({"amax_out": amax} if track else {}))
        if track:
            cat._amax = amax
        fused = self.vision_fusion.forward_nhwc(cat)                     # [M,Hs,Ws,96]
        Cf = fused.shape[-1]
        # grid_size = (nx, ny, nz); the BEV map is [grid[0] rows, grid[1] cols] (:248-250)
        out = ops.splat_soft(xy, fused.view(M, Hs * Ws, Cf), mask, grid[0], grid[1], self.min_weight,
                             want_nhwc=True, want_nchw=want_nchw)
        fa = engine.carried_amax(fused)
        if track and fa is not None and self.min_weight > 0:
            out["bev_nhwc"]._amax = fa
        ret = {"bev_densities": out["dens"], "bev_coords": xy}
        if want_nchw:
            ret["bev_features"] = out["bev_nchw"]
        return ret, out["bev_nhwc"]

    # ---- the reference's helper methods as callable entry points (same signatures and layouts) ----
    def _prepare_features_and_coords(self, x):
        """(depth [B,N,H,W], feats [B,N,F,H,W], p2p [B,N,4,4]) -> xyz [B,N,3,H,W], xyz_mask [B,N,1,H,W]
        bool, fused feats [B,N,C,H,W] (reference :131-173)."""
        require_eval(self)
        depth, feats, p2p = x
        B, N, F, H, W = feats.shape
        d = depth.reshape(B * N, H, W).float()
        m = p2p.reshape(B * N, 4, 4).float()
        xyz = ops.camera_to_world(d, m)                                       # [BN,3,H,W]
        _, z, mask = self.frustum(d, m)
        l0, l2 = self.z_proj[0], self.z_proj[2]
        w1, b1, w2, b2 = self._zmlp_weights(l0, l2)
        cat = ops.zmlp_concat(ops.nchw_to_nhwc(feats.reshape(B * N, F, H, W).float()), z, w1, b1, w2, b2)
        fused = ops.nhwc_to_nchw(self.vision_fusion.forward_nhwc(cat))
        return (xyz.view(B, N, 3, H, W), mask.view(B, N, 1, H, W).bool(),
                fused.view(B, N, fused.shape[1], H, W))

    def _points_to_voxels(self, points):
        """points [B,P,3] (LiDAR frame) -> fractional voxel coordinates [B,P,2] (reference :175-189)."""
        vox = [float(v) for v in self.voxel_size.tolist()]
        return ops.points_to_voxels(points, self.lidar2map.tolist(), vox[:2])

    def splat_soft(self, x):
        """(points_2d [B,P,2], points_features [B,F,P], grid_size (H, W)) -> (volume_features [B,F,H*W],
        volume_densities [B,H*W,1]) (reference :262-354, scatter_mode='mean')."""
        xy, feats, grid = x
        H, W = int(grid[0]), int(grid[1])
        B, F, P = feats.shape
        f = ops.nchw_to_nhwc(feats.reshape(B, F, 1, P).float()).view(B, P, F)
        Fp = F + (-F) % 4
        if Fp != F:
            f = torch.cat([f, f.new_zeros(B, P, Fp - F)], dim=-1).contiguous()
        out = ops.splat_soft(xy, f, None, H, W, self.min_weight, want_nhwc=False, want_nchw=True)
        return out["bev_nchw"][:, :F].reshape(B, F, H * W), out["dens"].view(B, H * W, 1)

    def _zmlp_weights(self, l0, l2):
        return self._cache.get(
            "z", [l0.weight, l0.bias, l2.weight, l2.bias],
            lambda: (l0.weight.detach().float().reshape(-1).contiguous(),
                     l0.bias.detach().float().contiguous(),
                     l2.weight.detach().float().contiguous(), l2.bias.detach().float().contiguous()))

    def forward_train(self, depth, feats_nhwc, p2p, want_nchw=True):
        """Train-mode splat as an autograd graph (reference :131-173, :191-260 under nn.Module.train(), stage 2):
        frustum -> z-MLP (two 1x1 convs over the height channel) -> concat -> fusion conv + BatchNorm (batch
        statistics) + ReLU -> bounds mask -> bilinear splat.  Gradients reach the image features, the z-MLP /
        fusion parameters and -- through the voxel coordinates and z -- the predicted depth."""
        from creste_public_b200 import autograd as ag
        M, Hs, Ws, F = feats_nhwc.shape
        rng, vox, grid = self._geom()
        xy, z, mask = ag.FrustumFn.apply(depth.contiguous().float(), p2p.contiguous().float(), rng, vox)
        l0, l2 = self.z_proj[0], self.z_proj[2]
        zf = z.view(M, Hs, Ws, 1)
        zf = ag.ChanAffineFn.apply(ag.Conv2dFn.apply(zf, l0.weight.view(l0.out_features, 1, 1, 1), 0, 0), None,
                                   l0.bias, True)
        zf = ag.ChanAffineFn.apply(ag.Conv2dFn.apply(zf, l2.weight.view(l2.out_features, l2.in_features, 1, 1), 0, 0),
                                   None, l2.bias, True)
        fused = self.vision_fusion.forward_nhwc(torch.cat([feats_nhwc, zf], dim=-1))       # BN batch statistics
        Cf = fused.shape[-1]
        bev_nhwc, dens = ag.SplatFn.apply(xy, fused.reshape(M, Hs * Ws, Cf), mask, grid[0], grid[1], self.min_weight)
        ret = {"bev_densities": dens, "bev_coords": xy}
        if want_nchw:
            ret["bev_features"] = ag.ToNCHW.apply(bev_nhwc)
        return ret, bev_nhwc

    def forward(self, x):
        assert len(x) >= 3, "Input must contain depth, features and camera projection matrix."
        if len(x) == 4 and self.training:
            raise NotImplementedError("movability masks are a training-only branch")
        depth, feats, p2p = x[:3]
        B, N, F, H, W = feats.shape
        assert N % self.NC == 0, f"Number of frames must be divisible by {self.NC}"
        f = ops.nchw_to_nhwc(feats.reshape(B * N, F, H, W).float())
        ret, _ = self.forward_nhwc(depth.reshape(B * N, H, W).float(), f,
                                   p2p.reshape(B * N, 4, 4).float())
        return ret


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/blocks/splat_projection.py")
