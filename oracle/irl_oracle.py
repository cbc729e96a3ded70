"""Stage-3 (counterfactual MaxEnt IRL) training-step oracle.  TEST / BENCH INFRASTRUCTURE.

Restates MaxEntIRLModel.training_step (reference creste/train_traversability.py:62-103) for the
head-only variant of SURVEY.md section 8(d) config 4: reward FCN (conv.py:88-161, train-mode
BatchNorm) on a given `input_view` -> MaxEntIRLLoss (loss_utils.py:1118-1259, incl. the
double-backward gradient penalty) -> backward -> Adam (lr 5e-4).  `reference_step` drives the
UNMODIFIED reference modules under the shims (build container only); `port_step` is the same
computation on plain torch CPU modules (travels to the GPU box: bench.py's cpu_baseline leg).
Only tests/, bench.py's cpu_baseline / --impl reference legs and smoke() may import this.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import net_oracle, synth

T_EXPERT = 50


def make_case(seed=3, B=2, H=8, W=16, T=T_EXPERT):
    """Seeded inputs of one head-only IRL step on an H x W reward grid (pooled cells)."""
    g = np.random.default_rng(9000 + seed)
    net = PortMSFCN()
    sd = synth.seeded_state_dict(net.state_dict(), seed)
    for k in sd:   # Xavier-scale conv weights as the reference initialises them
        if k.endswith("conv.weight"):
            fan_out = sd[k].shape[0] * sd[k].shape[2] * sd[k].shape[3]
            fan_in = sd[k].shape[1] * sd[k].shape[2] * sd[k].shape[3]
            sd[k] = sd[k] * float(np.sqrt(fan_in / 2.0) * np.sqrt(2.0 / (fan_in + fan_out)))
    expert = synth.expert_poses(B, T, 4 * H, 2 * W, seed)
    cfs = synth.counterfactuals(expert, every=2, shift=0.12 * 2 * W)
    fov = net_oracle.trapezoid_fov_mask(4 * H, 2 * W, 70, 70, 7 * W / 128.0, 200 * W / 128.0)
    return {
        "map_size": (H, W), "state_dict": sd,
        "input_view": torch.from_numpy(g.standard_normal((B, 40, H, W)).astype(np.float32)),
        "exp_svf": torch.from_numpy((g.random((B, H, W)) ** 4).astype(np.float32)),
        "expert": torch.from_numpy(expert), "cfs": cfs,
        "fov": torch.from_numpy(np.ascontiguousarray(fov)).unsqueeze(0).repeat(B, 1, 1),
    }


class FlatAdamTorch:
    """torch.optim.Adam with the reference's hyper-parameters (lr 5e-4, betas .9/.999)."""

    def __init__(self, params, lr=5e-4):
        self.opt = torch.optim.Adam(params, lr=lr)

    def zero_grad(self):
        self.opt.zero_grad()

    def step(self):
        self.opt.step()


def run_steps(net, loss_fn, case, steps=1, adam=FlatAdamTorch, device=None):
    """Generic driver: `net` maps input_view NCHW -> r [B,1,H,W]; `loss_fn.loss(tensor_dict)`
    follows the reference signature.  Returns numpy copies of everything a parity test needs."""
    dev = device or case["input_view"].device
    params = [p for p in net.parameters() if p.requires_grad]
    opt = adam(params)
    losses, metas, grads, r0 = [], {}, None, None
    for s in range(steps):
        iv = case["input_view"].to(dev).clone().requires_grad_(True)
        r = net(iv)
        td = {"outputs/exp_svf": case["exp_svf"].to(dev).clone(),
              "inputs/traversability_label": case["expert"].to(dev),
              "inputs/fov_mask": case["fov"].to(dev),
              "inputs/counterfactuals_label": case["cfs"],
              "outputs/traversability_preds": r, "outputs/input_view": iv}
        ld, md = loss_fn.loss(td)
        loss = ld["maxentirl_loss"]
        opt.zero_grad()
        loss.backward()
        if s == 0:
            grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in net.named_parameters()
                     if p.grad is not None}
            r0 = r.detach().cpu().numpy().copy()
        opt.step()
        losses.append(float(loss.detach().cpu()))
        for k, v in md.items():
            metas.setdefault(k, []).append(float(v.detach().cpu()))
    out = {"loss": np.array(losses, np.float32), "r": r0, "grads": grads,
           "params": {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}}
    out.update({k: np.array(v, np.float32) for k, v in metas.items()})
    return out


def reference_step(case, steps=1):
    """The unmodified reference MultiScaleFCN + MaxEntIRLLoss (build container only)."""
    from . import ref_harness as rh
    mods = rh.ref_modules()
    from omegaconf import OmegaConf
    cfgs = rh.compose_cfgs(map_size=case["map_size"])
    kw = cfgs["irl"]["traversability_head"]["net_kwargs"]["reward_cfg"]["net_kwargs"]
    net = mods["conv"].MultiScaleFCN(OmegaConf.create(kw))
    net.load_state_dict(case["state_dict"])
    net.train()
    loss_fn = mods["loss_utils"].MaxEntIRLLoss(OmegaConf.create(cfgs["irl"]["loss"][0]))
    return run_steps(net, loss_fn, case, steps)


# ------------------------------------------------------------------------- plain-torch port
class _ConvLayer(nn.Sequential):
    def __init__(self, cin, cout, k, bn, relu=True):
        super().__init__()
        self.add_module("conv", nn.Conv2d(cin, cout, k, padding=k // 2, bias=False))
        if bn:
            self.add_module("norm", nn.BatchNorm2d(cout))
        if relu:
            self.add_module("relu", nn.ReLU(inplace=True))


class PortMSFCN(nn.Module):
    """conv.py:88-161 with the dims of terrainnet_maxentirlcf_msfcn_sam2dynsemelev.yaml:36-60;
    parameter names equal the reference's (state dicts are interchangeable)."""

    def __init__(self):
        super().__init__()
        self.prepool = nn.Sequential(_ConvLayer(40, 64, 5, True), _ConvLayer(64, 32, 3, True))
        self.skip = nn.Sequential(_ConvLayer(32, 32, 3, True), _ConvLayer(32, 16, 1, True))
        self.trunk = nn.Sequential(nn.MaxPool2d(2, 2), _ConvLayer(32, 32, 3, False),
                                   nn.BatchNorm2d(32), nn.ReLU(inplace=True),
                                   _ConvLayer(32, 32, 1, False), nn.BatchNorm2d(32),
                                   nn.ReLU(inplace=True),
                                   nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
        self.postpool = nn.Sequential(_ConvLayer(48, 1, 1, True))

    def forward(self, x):
        x = self.prepool(x)
        skip = self.skip(x)
        x = self.trunk(x)
        return self.postpool(torch.cat([x, skip], dim=1))


class PortLoss:
    """MaxEntIRLLoss.loss (loss_utils.py:1118-1259) restated on torch CPU."""

    def __init__(self, map_sz, map_ds=2, alpha=0.5, reward_weight=0.01, maxent_weight=1.0,
                 use_fov_mask=True):
        self.map_sz, self.map_ds, self.alpha = tuple(map_sz), map_ds, alpha
        self.reward_weight, self.maxent_weight, self.use_fov_mask = reward_weight, maxent_weight, use_fov_mask

    def visitation(self, xy):
        xy = xy if xy.ndim == 3 else xy[:, :, :2, 2]
        xy = xy / self.map_ds
        H, W = self.map_sz
        B = xy.shape[0]
        s, e = xy[:, :-1], xy[:, 1:]
        ms = torch.ceil(torch.norm(e - s, dim=-1)).long().max().item()
        t = torch.linspace(0, 1, ms).view(1, 1, -1, 1)
        pts = (s.unsqueeze(2) + t * (e - s).unsqueeze(2)).view(B, -1, 2)
        pts = torch.cat([pts, xy[:, -1:]], dim=1)
        idx = pts[:, :, 0].clamp(0, H - 1).long() * W + pts[:, :, 1].clamp(0, W - 1).long()
        cnt = torch.zeros(B, H * W, dtype=torch.float32)
        cnt.scatter_add_(1, idx, torch.ones_like(idx, dtype=torch.float32))
        cnt[cnt > 1] = 1
        return cnt.view(B, H, W)

    def loss(self, td):
        exp_svf, gt, fov = td["outputs/exp_svf"], td["inputs/traversability_label"], td["inputs/fov_mask"]
        r = td["outputs/traversability_preds"].squeeze(1)
        iv = td["outputs/input_view"]
        _, Ho, Wo = fov.shape
        _, H, W = exp_svf.shape
        fm = F.interpolate(fov.unsqueeze(1).byte(), size=(Ho // 2, Wo // 2), mode="nearest")
        fm = fm[..., :H, :W].squeeze(1).bool()
        svf = self.visitation(gt)
        if self.use_fov_mask:
            svf, exp_svf = svf * fm.float(), exp_svf * fm.float()
        svf = svf / (svf.sum(dim=(1, 2), keepdim=True) + 1e-5)
        exp_svf = exp_svf / (exp_svf.sum(dim=(1, 2), keepdim=True) + 1e-5)
        cf_tot, exp_tot = torch.zeros_like(svf), exp_svf.clone()
        for i, cf in enumerate(td["inputs/counterfactuals_label"]):
            if cf is None:
                continue
            bad = cf["trajectories"][cf["rank"] > 0]
            if bad.shape[0] == 0:
                continue
            c = self.visitation(torch.from_numpy(bad)).sum(0)
            c = c / (c.sum() + 1e-5)
            exp_svf[i] = self.alpha * c + (1 - self.alpha) * exp_svf[i]
            cf_tot[i] = c
        if self.use_fov_mask:
            r = r * fm.float()
        e_r, s_r = (exp_svf * r).sum(dim=(1, 2)).mean(), (svf * r).sum(dim=(1, 2)).mean()
        pen = torch.tensor(0.0)
        if r.requires_grad and self.reward_weight > 0:
            g = torch.autograd.grad(r.sum(), iv, create_graph=True, retain_graph=True)[0]
            pen = ((g.norm(2, dim=1) - 1) ** 2).mean()
        loss = self.maxent_weight * (e_r - s_r) + self.reward_weight * pen
        with torch.no_grad():
            cr, orr = (cf_tot * r).sum(dim=(1, 2)), (exp_tot * r).sum(dim=(1, 2))
            v = cr != 0
        return {"maxentirl_loss": loss}, {"reward_penalty": self.reward_weight * pen,
                                          "mean_expected_svf_rewards": e_r, "mean_svf_rewards": s_r,
                                          "sum_cf_rewards": cr[v].sum(), "sum_opt_rewards": orr[v].sum()}


def pool_argmax(x_nchw):
    """Per 2x2/2 window of [B,C,H,W]: (index 0..3 of the maximum in row-major window order, top-1 minus top-2)."""
    B, Cc, H, W = x_nchw.shape
    w = x_nchw.reshape(B, Cc, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, Cc, H // 2, W // 2, 4)
    top = w.topk(2, dim=-1)
    return top.indices[..., 0], top.values[..., 0] - top.values[..., 1]


def port_step(case, steps=1, dtype=torch.float32, pool_hook=None, relu_hook=None):
    """dtype=float64 gives the 'exact' answer used as the yardstick of fp32 conditioning.
    pool_hook(x) -> x' | None is called on the max-pool's input (NCHW), relu_hook(i, u) -> u' | None on the input of
    the i-th ReLU in execution order: parity tests use them to read the oracle's DISCRETE decisions (pool routing,
    ReLU masks) and to break near-ties (two window values, or a pre-activation and zero, closer than fp32 rounding
    noise) the way the implementation under test broke them, so that the comparison of the gradients is not a
    comparison of coin flips."""
    net = PortMSFCN()
    net.load_state_dict(case["state_dict"])
    net.train()
    if pool_hook is not None:
        net.trunk[0].register_forward_pre_hook(lambda mod, args: pool_hook(args[0]))
    if relu_hook is not None:
        count = [0]

        def pre(mod, args):
            i = count[0]
            count[0] += 1
            return relu_hook(i, args[0])

        for m in net.modules():
            if isinstance(m, nn.ReLU):
                m.register_forward_pre_hook(pre)
    if dtype != torch.float32:
        net = net.to(dtype)
        case = dict(case)
        case["input_view"] = case["input_view"].to(dtype)
        case["exp_svf"] = case["exp_svf"].to(dtype)
    return run_steps(net, PortLoss(case["map_size"]), case, steps)


class PortHeadStep:
    """Head-only stage-3 step on the host (primary variant), the CPU
    counterpart of creste.train_traversability.HeadStep: cat + 2x2 max-pool + top-half crop
    (vin.py:104-116) -> reward FCN (train mode) -> value iteration + SVF + rollout (the C oracle)
    -> MaxEntIRLLoss -> backward -> Adam."""

    def __init__(self, state_dict, map_size, action_horizon=50, lr=5e-4):
        self.net = PortMSFCN()
        self.net.load_state_dict(state_dict)
        self.net.train()
        self.loss = PortLoss(map_size)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr)
        self.map_size, self.T = tuple(map_size), action_horizon
        H, W = self.map_size
        self.fov_model = net_oracle.trapezoid_fov_mask(2 * H, W, 70, 70, 0, 100)[:H, :W]

    def __call__(self, feat_nchw, expert, fov_mask, cfs):
        from . import c_oracle as co
        x = torch.cat(feat_nchw, dim=1)
        iv = F.max_pool2d(x, 2, 2)
        iv = iv[:, :, : iv.shape[2] // 2].detach().requires_grad_(True)
        r = self.net(iv)
        v, q, pi, K = co.vi_solve(r.detach().numpy())
        svf, states, grid = co.svf(pi, expert[:, :, :2, 2].numpy().copy(), self.fov_model, self.T, 2,
                                   True, 0.005, False)
        td = {"outputs/exp_svf": torch.from_numpy(svf), "inputs/traversability_label": expert,
              "inputs/fov_mask": fov_mask, "inputs/counterfactuals_label": cfs,
              "outputs/traversability_preds": r, "outputs/input_view": iv}
        ld, md = self.loss.loss(td)
        self.opt.zero_grad()
        ld["maxentirl_loss"].backward()
        self.opt.step()
        return float(ld["maxentirl_loss"].detach()), {"r": r.detach(), "K": K, "exp_svf": svf, "v": v,
                                             "states": states, **{k: float(v_.detach()) for k, v_ in md.items()}}


def head_inputs(B, Hm, Wm, seed=0, T=T_EXPERT):
    return synth.head_inputs(B, Hm, Wm, seed, T)
