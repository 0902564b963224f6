"""oracle/dp_oracle.py -- TEST INFRASTRUCTURE (never imported by the product path).

CPU restatement, in plain fp32 torch, of the Diffusion-Policy training path of the reference
(SURVEY.md section 8 rows a11, a12):
  * `PCDObsEncoder`              src/models/components/diffusion_policy/vision/pcd_obs_encoder.py:14-296
  * `ConditionalUnet1D` & blocks src/models/components/diffusion_policy/diffusion/conditional_unet1d.py:17-297,
                                 conv1d_components.py:8-45, positional_embedding.py:7-19
  * `DiffusionUnetImagePolicy.compute_loss`   .../diffusion_unet_image_policy.py:233-313
  * `LinearNormalizer`           src/utils/diffusion_policy/normalizer.py:14-78,195-300
  * `LowdimMaskGenerator`        .../diffusion/mask_generator.py:41-105
  * `DDPMScheduler.add_noise`    THIRD PARTY, absent from /root/reference: diffusers 0.29.0
                                 (requirements: `diffusers==0.29.0`), schedulers/scheduling_ddpm.py
                                 -- `betas_for_alpha_bar` (cosine, max_beta 0.999) and
                                 `add_noise` = sqrt(acp[t]) x + sqrt(1 - acp[t]) eps, restated from
                                 the published algorithm (Nichol & Dhariwal 2021, eq. 17).
Module tree and `state_dict` keys equal the reference's, so its checkpoints load unchanged.

Pinning: tests/golden/dp_*.npz are produced by the REFERENCE's own modules
(oracle/gen_golden_dp.py); tests/test_dp_oracle_cpu.py holds this file to them.  The scheduler is
the one piece whose reference implementation cannot be run here: its parity is unpinned (the
fixture generator injects this file's `DDPMSchedule` into the reference policy).
Set-abstraction indices come from the C pointops oracle (oracle/pointops_oracle.c).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .act_oracle import grouping_with_xyz, oracle_fps, oracle_knn


# ---- diffusers 0.29.0 scheduling_ddpm.py (restated) ---------------------------------------------
def betas_for_alpha_bar(n, max_beta=0.999):
    def alpha_bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = []
    for i in range(n):
        t1, t2 = i / n, (i + 1) / n
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class DDPMSchedule:
    """`DDPMScheduler(num_train_timesteps, beta_start, beta_end, beta_schedule, prediction_type, ...)`
    as configured by configs/model/maniskill2_diffusion_policy_model.yaml:29-38."""

    def __init__(self, num_train_timesteps=100, beta_start=0.0001, beta_end=0.02, beta_schedule="squaredcos_cap_v2",
                 clip_sample=True, prediction_type="epsilon", variance_type="fixed_small", **_):
        if beta_schedule == "squaredcos_cap_v2":
            betas = betas_for_alpha_bar(num_train_timesteps)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                           clip_sample=clip_sample, clip_sample_range=1.0, variance_type=variance_type)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    # diffusers 0.29.0 scheduling_ddpm.py: set_timesteps ("leading" spacing, the default) / step /
    # _get_variance ("fixed_small"), restated from the published sampler (Ho et al. 2020, eq. 6-7, 11)
    def set_timesteps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        self.timesteps = torch.arange(num_inference_steps - 1, -1, -1) * ratio

    def step(self, model_output, t, sample, generator=None, noise=None, **_):
        """-> object with `.prev_sample` (the attribute the reference's sampling loop reads,
        diffusion_unet_image_policy.py:138-140)."""
        t = int(t)
        n = self.num_inference_steps or self.config.num_train_timesteps
        prev_t = t - self.config.num_train_timesteps // n
        acp = self.alphas_cumprod.to(sample.device)
        a_t = acp[t]
        a_prev = acp[prev_t] if prev_t >= 0 else torch.tensor(1.0, device=sample.device)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        if self.config.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        else:
            x0 = model_output
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        prev = (a_prev ** 0.5 * cur_b) / b_t * x0 + cur_a ** 0.5 * b_prev / b_t * sample
        if t > 0:
            var = torch.clamp(b_prev / b_t * cur_b, min=1e-20)
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                    dtype=model_output.dtype)
            prev = prev + var ** 0.5 * noise
        return _Cfg(prev_sample=prev)

    def add_noise(self, x, noise, timesteps):
        acp = self.alphas_cumprod.to(device=x.device, dtype=x.dtype)
        a = acp[timesteps] ** 0.5
        s = (1 - acp[timesteps]) ** 0.5
        while a.dim() < x.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * x + s * noise


# ---- normalizer.py -----------------------------------------------------------------------------
class OracleLinearNormalizer(nn.Module):
    def __init__(self):
        super().__init__()
        self.params_dict = nn.ParameterDict()

    def set_field(self, key, scale, offset):
        self.params_dict[key] = nn.ParameterDict({"scale": nn.Parameter(scale.clone(), requires_grad=False),
                                                  "offset": nn.Parameter(offset.clone(), requires_grad=False)})

    def normalize_field(self, key, x):
        p = self.params_dict[key]
        shape = x.shape
        return (x.reshape(-1, p["scale"].shape[0]) * p["scale"] + p["offset"]).reshape(shape)

    def _load_from_state_dict(self, state_dict, prefix, *args):  # dict_of_tensor_mixin.py:15-47
        for k, v in state_dict.items():
            if k.startswith(prefix + "params_dict."):
                field, name = k[len(prefix + "params_dict."):].split(".", 1)
                if name in ("scale", "offset"):
                    if field not in self.params_dict:
                        self.params_dict[field] = nn.ParameterDict()
                    self.params_dict[field][name] = nn.Parameter(v.clone(), requires_grad=False)


# ---- conv1d_components.py / positional_embedding.py / conditional_unet1d.py -------------------
class _Wrap(nn.Module):
    def __init__(self, conv):
        super().__init__()
        self.conv = conv

    def forward(self, x):
        return self.conv(x)


class OracleConv1dBlock(nn.Module):
    def __init__(self, cin, cout, k, n_groups=8):
        super().__init__()
        self.block = nn.Sequential(nn.Conv1d(cin, cout, k, padding=k // 2), nn.GroupNorm(n_groups, cout), nn.Mish())

    def forward(self, x):
        return self.block(x)


class OracleResBlock(nn.Module):
    def __init__(self, cin, cout, cond_dim, kernel_size=3, n_groups=8, cond_predict_scale=False):
        super().__init__()
        self.blocks = nn.ModuleList([OracleConv1dBlock(cin, cout, kernel_size, n_groups),
                                     OracleConv1dBlock(cout, cout, kernel_size, n_groups)])
        self.cond_predict_scale, self.out_channels = cond_predict_scale, cout
        self.cond_encoder = nn.Sequential(nn.Mish(), nn.Linear(cond_dim, cout * 2 if cond_predict_scale else cout),
                                          nn.Identity())
        self.residual_conv = nn.Conv1d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x, cond):
        out = self.blocks[0](x)
        embed = self.cond_encoder(cond).unsqueeze(-1)
        if self.cond_predict_scale:
            embed = embed.reshape(embed.shape[0], 2, self.out_channels, 1)
            out = embed[:, 0] * out + embed[:, 1]
        else:
            out = out + embed
        return self.blocks[1](out) + self.residual_conv(x)


class _SinPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        half = self.dim // 2
        e = torch.exp(torch.arange(half, device=x.device) * -(math.log(10000) / (half - 1)))
        e = x[:, None] * e[None, :]
        return torch.cat((e.sin(), e.cos()), dim=-1)


class OracleConditionalUnet1D(nn.Module):
    def __init__(self, input_dim, local_cond_dim=None, global_cond_dim=None, diffusion_step_embed_dim=256,
                 down_dims=(256, 512, 1024), kernel_size=3, n_groups=8, cond_predict_scale=False):
        super().__init__()
        assert local_cond_dim is None  # never set on the reference's training path (diffusion_unet_image_policy.py:72)
        all_dims = [input_dim] + list(down_dims)
        dsed = diffusion_step_embed_dim
        self.diffusion_step_encoder = nn.Sequential(_SinPosEmb(dsed), nn.Linear(dsed, dsed * 4), nn.Mish(),
                                                    nn.Linear(dsed * 4, dsed))
        cond_dim = dsed + (global_cond_dim or 0)
        in_out = list(zip(all_dims[:-1], all_dims[1:]))
        kw = dict(cond_dim=cond_dim, kernel_size=kernel_size, n_groups=n_groups, cond_predict_scale=cond_predict_scale)
        mid = all_dims[-1]
        self.local_cond_encoder = None
        self.mid_modules = nn.ModuleList([OracleResBlock(mid, mid, **kw), OracleResBlock(mid, mid, **kw)])
        self.down_modules = nn.ModuleList()
        for ind, (di, do) in enumerate(in_out):
            last = ind >= len(in_out) - 1
            self.down_modules.append(nn.ModuleList([OracleResBlock(di, do, **kw), OracleResBlock(do, do, **kw),
                                                    _Wrap(nn.Conv1d(do, do, 3, 2, 1)) if not last else nn.Identity()]))
        self.up_modules = nn.ModuleList()
        for ind, (di, do) in enumerate(reversed(in_out[1:])):
            last = ind >= len(in_out) - 1
            self.up_modules.append(nn.ModuleList([OracleResBlock(do * 2, di, **kw), OracleResBlock(di, di, **kw),
                                                  _Wrap(nn.ConvTranspose1d(di, di, 4, 2, 1)) if not last else nn.Identity()]))
        self.final_conv = nn.Sequential(OracleConv1dBlock(down_dims[0], down_dims[0], kernel_size),
                                        nn.Conv1d(down_dims[0], input_dim, 1))

    def forward(self, sample, timestep, local_cond=None, global_cond=None):
        x = sample.permute(0, 2, 1)
        g = self.diffusion_step_encoder(timestep.expand(sample.shape[0]))
        if global_cond is not None:
            g = torch.cat([g, global_cond], dim=-1)
        h = []
        for r1, r2, down in self.down_modules:
            x = r2(r1(x, g), g)
            h.append(x)
            x = down(x)
        for m in self.mid_modules:
            x = m(x, g)
        for r1, r2, up in self.up_modules:
            x = up(r2(r1(torch.cat((x, h.pop()), dim=1), g), g))
        return self.final_conv(x).permute(0, 2, 1)


# ---- pcd_obs_encoder.py --------------------------------------------------------------------------
class OraclePCDObsEncoder(nn.Module):
    def __init__(self, shape_meta, pcd_model, share_pcd_model=True, n_obs_step=2, pcd_nsample=16, pcd_npoints=1024,
                 use_mask=False, bg_ratio=0.0, pcd_hidden_dim=128, projector_layers=2,
                 projector_channels=(128, 128, 128), pre_sample=False, in_channel=6, **_):
        super().__init__()
        assert share_pcd_model
        self.use_mask, self.bg_ratio, self.pre_sample = use_mask, bg_ratio, pre_sample
        self.key_model_map = nn.ModuleDict({"pcd": pcd_model})
        self.shape_meta, self.n_obs_step = shape_meta, n_obs_step
        self.pcd_keys = sorted(k for k, a in shape_meta["obs"].items() if a.get("type", "low_dim") == "pcd")
        self.low_dim_keys = sorted(k for k, a in shape_meta["obs"].items() if a.get("type", "low_dim") == "low_dim")
        self.pcd_nsample, self.pcd_npoints = pcd_nsample, pcd_npoints
        if not pre_sample:
            self.linear = nn.Linear(3 + pcd_model.num_channels, pcd_hidden_dim, bias=False)
            self.bn = nn.BatchNorm1d(pcd_hidden_dim)
        else:  # pcd_obs_encoder.py:91-93
            self.linear = nn.Linear(3 + in_channel, in_channel, bias=False)
            self.bn = nn.BatchNorm1d(in_channel)
        proj = []
        for i in range(projector_layers):
            cin = pcd_model.num_channels if (i == 0 and pre_sample) else pcd_hidden_dim  # :101-110
            proj += [nn.Conv1d(cin, projector_channels[i], 1), nn.BatchNorm1d(projector_channels[i]), nn.ReLU()]
        proj += [nn.MaxPool1d(pcd_npoints), nn.Conv1d(projector_channels[i], projector_channels[i + 1], 1),
                 nn.BatchNorm1d(projector_channels[i + 1])]
        self.projector = nn.Sequential(*proj)
        self.projector_channels = list(projector_channels)
        self._dummy_variable = nn.Parameter(torch.empty(0))  # module_attr_mixin.py:7-9 (state_dict key)

    def output_dim(self):
        return self.projector_channels[-1] + sum(int(self.shape_meta["obs"][k]["shape"][0]) for k in self.low_dim_keys)

    def pcd_sampling(self, p, x, o, mask):
        """pcd_obs_encoder.py:123-198; the masked branch is the one of ACT (act_oracle.OracleACTPCD.pcd_sampling)."""
        b = o.shape[0]
        n_o = torch.arange(1, b + 1, dtype=torch.int32) * self.pcd_npoints
        if not self.use_mask or mask is None:
            idx = oracle_fps(p, o, n_o)
        else:
            n_bg = int(self.pcd_npoints * self.bg_ratio)
            ar = torch.arange(1, b + 1, dtype=torch.int32)
            ends = o.long()
            fg_o = torch.cumsum(mask.long(), 0)[ends - 1].int()
            idx = oracle_fps(p[mask].contiguous(), fg_o, ar * (self.pcd_npoints - n_bg) if self.bg_ratio > 0.0 else n_o)
            if self.bg_ratio > 0.0:
                bg_o = torch.cumsum((~mask).long(), 0)[ends - 1].int()
                idx = torch.cat([idx, oracle_fps(p[~mask].contiguous(), bg_o, ar * n_bg)], 0)
        n_p = p[idx.long(), :]
        kidx = oracle_knn(self.pcd_nsample, p, o, n_p, n_o)
        g = grouping_with_xyz(kidx, x, p, n_p)
        y = F.relu(self.bn(self.linear(g).transpose(1, 2).contiguous())).max(dim=-1).values  # (m, c)
        return n_p, y, n_o, idx

    def encode_pcd(self, pcd):
        mask = pcd.get("mask") if self.use_mask else None
        b = pcd["offset"].shape[0]
        if self.pre_sample:  # pcd_obs_encoder.py:201-218
            coord, feats, off, idx = self.pcd_sampling(pcd["coord"], pcd["feat"], pcd["offset"], mask)
            y = self.key_model_map["pcd"](dict(pcd, coord=coord, feat=feats, offset=off,
                                               grid_coord=pcd["grid_coord"][idx.long()]))
        else:
            feats = self.key_model_map["pcd"](pcd)
            _, y, _, _ = self.pcd_sampling(pcd["coord"], feats, pcd["offset"], mask)
        x = y.view(b, self.pcd_npoints, -1).permute(0, 2, 1)
        return self.projector(x).squeeze(-1)

    def forward(self, obs):
        feats = [self.encode_pcd(obs[k]) for k in self.pcd_keys] + [obs[k] for k in self.low_dim_keys]
        return torch.cat(feats, dim=-1)


# ---- diffusion_unet_image_policy.py --------------------------------------------------------------
class OracleDiffusionPolicy(nn.Module):
    def __init__(self, shape_meta, noise_scheduler, obs_encoder, horizon, n_action_steps, n_obs_steps,
                 num_inference_steps=None, obs_as_global_cond=True, diffusion_step_embed_dim=256,
                 down_dims=(256, 512, 1024), kernel_size=5, n_groups=8, cond_predict_scale=True, **_):
        super().__init__()
        assert obs_as_global_cond  # compute_loss raises otherwise (diffusion_unet_image_policy.py:259-260)
        self.action_dim = int(shape_meta["action"]["shape"][0])
        gdim = obs_encoder.output_dim() * n_obs_steps
        goal = shape_meta.get("goal")
        if goal is not None:
            gdim += int(goal["task_emb"]["shape"][0])
        self.obs_encoder = obs_encoder
        self.model = OracleConditionalUnet1D(self.action_dim, None, gdim, diffusion_step_embed_dim, down_dims,
                                             kernel_size, n_groups, cond_predict_scale)
        self.noise_scheduler = noise_scheduler
        self.normalizer = OracleLinearNormalizer()
        self.horizon, self.n_action_steps, self.n_obs_steps = horizon, n_action_steps, n_obs_steps
        self.num_inference_steps = num_inference_steps or noise_scheduler.config.num_train_timesteps
        self._dummy_variable = nn.Parameter(torch.empty(0))  # ModuleAttrMixin of the policy / mask generator
        self.mask_generator = nn.Module()
        self.mask_generator._dummy_variable = nn.Parameter(torch.empty(0))

    def compute_loss(self, batch):
        obs = dict(batch["obs"])
        pcds = obs.pop("pcds", None)
        nobs = {k: self.normalizer.normalize_field(k, v) for k, v in obs.items()}
        nact = self.normalizer.normalize_field("action", batch["action"])
        bs = nact.shape[0]
        this = {k: v[:, : self.n_obs_steps].reshape(-1, *v.shape[2:]) for k, v in nobs.items()}
        if pcds is not None:
            this["pcds"] = pcds
        gcond = self.obs_encoder(this).reshape(bs, -1)
        if "goal" in batch and "task_emb" in batch["goal"]:
            gcond = torch.cat([gcond, batch["goal"]["task_emb"]], dim=-1)
        # LowdimMaskGenerator(obs_dim=0, action_visible=False): the mask is all False
        # (mask_generator.py:70-105), so nothing is in-painted and every element is in the loss
        noise = batch["_noise"] if "_noise" in batch else torch.randn(nact.shape)
        t = batch["_timesteps"] if "_timesteps" in batch else torch.randint(
            0, self.noise_scheduler.config.num_train_timesteps, (bs,)).long()
        noisy = self.noise_scheduler.add_noise(nact, noise, t)
        pred = self.model(noisy, t, global_cond=gcond)
        target = noise if self.noise_scheduler.config.prediction_type == "epsilon" else nact
        loss = F.mse_loss(pred, target, reduction="none").reshape(bs, -1).mean(1).mean()
        return dict(loss=loss, pred=pred)

    forward = compute_loss

    # diffusion_unet_image_policy.py:106-231 (conditional_sample + predict_action), observations as global
    # conditioning: no in-painting (the condition mask is all False).  `noises` = [x_T, eps_1, eps_2, ...]
    # replaces the sampler's draws (trajectory init, then one per step with t > 0).
    @torch.no_grad()
    def predict_action(self, obs_dict, noises=None):
        obs = dict(obs_dict["obs"])
        pcds = obs.pop("pcds", None)
        nobs = {k: self.normalizer.normalize_field(k, v) for k, v in obs.items()}
        bs = next(iter(nobs.values())).shape[0]
        this = {k: v[:, : self.n_obs_steps].reshape(-1, *v.shape[2:]) for k, v in nobs.items()}
        if pcds is not None:
            this["pcds"] = pcds
        gcond = self.obs_encoder(this).reshape(bs, -1)
        if "goal" in obs_dict and "task_emb" in obs_dict["goal"]:
            gcond = torch.cat([gcond, obs_dict["goal"]["task_emb"]], dim=-1)
        shape = (bs, self.horizon, self.action_dim)
        it = iter(noises) if noises is not None else None
        traj = next(it) if it is not None else torch.randn(shape)
        sch = self.noise_scheduler
        sch.set_timesteps(self.num_inference_steps)
        for t in sch.timesteps:
            out = self.model(traj, t, global_cond=gcond)
            traj = sch.step(out, t, traj, noise=(next(it) if (it is not None and int(t) > 0) else None)).prev_sample
        p = self.normalizer.params_dict["action"]
        action_pred = ((traj.reshape(-1, self.action_dim) - p["offset"]) / p["scale"]).reshape(shape)
        start = self.n_obs_steps - 1
        return {"action": action_pred[:, start:start + self.n_action_steps], "action_pred": action_pred}


def build_oracle_dp(cfg: dict):
    """cfg keys mirror scratch_pointnet_pcd.yaml + maniskill2_diffusion_policy_model.yaml."""
    from .act_oracle import OraclePointNet

    shape_meta = {"obs": {"pcds": {"shape": [6], "type": "pcd"}, "qpos": {"shape": [cfg["qpos_dim"]], "type": "low_dim"}},
                  "action": {"shape": [cfg["action_dim"]]}, "goal": None}
    if cfg.get("goal_dim", 0):
        shape_meta["goal"] = {"task_emb": {"shape": [cfg["goal_dim"]]}}
    enc = OraclePCDObsEncoder(shape_meta, OraclePointNet(6, cfg["backbone_classes"]), n_obs_step=cfg["n_obs_steps"],
                              pcd_nsample=cfg["pcd_nsample"], pcd_npoints=cfg["pcd_npoints"],
                              pcd_hidden_dim=cfg["pcd_hidden_dim"], projector_layers=cfg["projector_layers"],
                              projector_channels=cfg["projector_channels"])
    return OracleDiffusionPolicy(shape_meta, DDPMSchedule(num_train_timesteps=cfg.get("num_train_timesteps", 100)), enc,
                                 horizon=cfg["horizon"], n_action_steps=cfg.get("n_action_steps", 8),
                                 n_obs_steps=cfg["n_obs_steps"], num_inference_steps=cfg.get("num_inference_steps", None),
                                 diffusion_step_embed_dim=cfg["diffusion_step_embed_dim"],
                                 down_dims=cfg["down_dims"], kernel_size=cfg["kernel_size"], n_groups=cfg["n_groups"],
                                 cond_predict_scale=cfg.get("cond_predict_scale", True))
