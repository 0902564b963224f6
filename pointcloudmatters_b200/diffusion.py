"""Diffusion Policy on point-cloud observations -- host-side mirror of the reference's
`src/models/components/diffusion_policy/`:
  * `PCDObsEncoder`              vision/pcd_obs_encoder.py:14-296          
  * `ConditionalUnet1D` & blocks diffusion/conditional_unet1d.py:17-297, conv1d_components.py:8-45,
                                 positional_embedding.py:7-19              (row a12)
  * `DiffusionUnetImagePolicy`   diffusion_unet_image_policy.py:22-313 (training: `compute_loss`)
  * `LinearNormalizer`           src/utils/diffusion_policy/normalizer.py:14-300
  * `DDPMScheduler`              diffusers 0.29.0 (third party): the closed-form forward process only
Same constructor kwargs, module tree and `state_dict` keys (reference checkpoints load unchanged),
same batch contract (`{obs: {qpos, pcds}, action[, goal: {task_emb}]}` -> `{loss}`).

What changed underneath (B200-first):
  * activations of the denoiser are channel-last (B, T, C); every Conv1d / ConvTranspose1d / Linear is
    a tcgen05 GEMM over rows = B*T on the weight in its torch layout (functional_unet.py);
    GroupNorm + Mish + FiLM + residual add is one kernel; Mish of the conditioning vector is
    computed once per step instead of once per residual block (16x);
  * the set-abstraction head is the fused operator shared with ACT; FPS / kNN are sync-free when
    `pcds["n_max"]` is given.
CUDA only -- no CPU fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as PF
from . import functional_unet as UF
from . import pointops


# ---- conv1d_components.py ------------------------------------------------------------------------
class Conv1dCL(nn.Conv1d):
    """nn.Conv1d parameters; forward on channel-last (B, T, C) activations."""

    def forward(self, x):
        return UF.conv1d_cl(x, self.weight, self.bias, self.stride[0], self.padding[0])


class ConvTranspose1dCL(nn.ConvTranspose1d):
    def forward(self, x):
        return UF.conv_transpose1d_cl(x, self.weight, self.bias, self.stride[0], self.padding[0])


class Downsample1d(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = Conv1dCL(dim, dim, 3, 2, 1)

    def forward(self, x):
        return self.conv(x)


class Upsample1d(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = ConvTranspose1dCL(dim, dim, 4, 2, 1)

    def forward(self, x):
        return self.conv(x)


class Conv1dBlock(nn.Module):
    """Conv1d -> GroupNorm -> Mish (conv1d_components.py:25-44); the norm + activation (+ FiLM, + residual)
    run as one kernel."""

    def __init__(self, inp_channels, out_channels, kernel_size, n_groups=8):
        super().__init__()
        self.block = nn.Sequential(Conv1dCL(inp_channels, out_channels, kernel_size, padding=kernel_size // 2),
                                   nn.GroupNorm(n_groups, out_channels), nn.Mish())

    def forward(self, x, film=None, res=None):
        return UF.groupnorm_mish(self.block[0](x), self.block[1], film, res)


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        half = self.dim // 2
        e = torch.exp(torch.arange(half, device=x.device) * -(math.log(10000) / (half - 1)))
        e = x[:, None] * e[None, :]
        return torch.cat((e.sin(), e.cos()), dim=-1)


def _linear_padded(x, weight, bias):
    """PF.linear with the reduction dimension zero-padded to the TMA granularity (cond_dim = 128 + 2*137 = 402
    in the reference configuration): keeps the FiLM projections on the tensor-core path."""
    K_ = weight.shape[1]
    if K_ % 8 == 0:
        return PF.linear(x, weight, bias)
    pad = -(-K_ // 8) * 8 - K_
    if x.shape[-1] == K_:
        x = F.pad(x, (0, pad))
    return PF.linear(x, F.pad(weight, (0, pad)), bias)


class ConditionalResidualBlock1D(nn.Module):
    def __init__(self, in_channels, out_channels, cond_dim, kernel_size=3, n_groups=8, cond_predict_scale=False):
        super().__init__()
        self.blocks = nn.ModuleList([Conv1dBlock(in_channels, out_channels, kernel_size, n_groups=n_groups),
                                     Conv1dBlock(out_channels, out_channels, kernel_size, n_groups=n_groups)])
        cond_channels = out_channels * 2 if cond_predict_scale else out_channels
        self.cond_predict_scale, self.out_channels = cond_predict_scale, out_channels
        # index 2 is einops' parameter-free Rearrange in the reference
        self.cond_encoder = nn.Sequential(nn.Mish(), nn.Linear(cond_dim, cond_channels), nn.Identity())
        self.residual_conv = Conv1dCL(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()

    def forward(self, x, mish_cond):
        """x (B, T, Cin) channel-last; mish_cond = Mish(cond) (B, cond_dim[+pad]), shared by all blocks."""
        lin = self.cond_encoder[1]
        embed = _linear_padded(mish_cond, lin.weight, lin.bias)
        if not self.cond_predict_scale:  # out + embed (conditional_unet1d.py:78-79): scale 1
            embed = torch.cat([torch.ones_like(embed), embed], dim=-1)
        out = self.blocks[0](x, film=embed)
        return self.blocks[1](out, res=self.residual_conv(x))


class ConditionalUnet1D(nn.Module):
    def __init__(self, input_dim, local_cond_dim=None, global_cond_dim=None, diffusion_step_embed_dim=256,
                 down_dims=(256, 512, 1024), kernel_size=3, n_groups=8, cond_predict_scale=False):
        super().__init__()
        if local_cond_dim is not None:
            raise NotImplementedError("local conditioning is never enabled on the reference's training path "
                                      "(diffusion_unet_image_policy.py:72)")
        all_dims = [input_dim] + list(down_dims)
        dsed = diffusion_step_embed_dim
        self.diffusion_step_encoder = nn.Sequential(SinusoidalPosEmb(dsed), nn.Linear(dsed, dsed * 4), nn.Mish(),
                                                    nn.Linear(dsed * 4, dsed))
        cond_dim = dsed + (global_cond_dim or 0)
        self.cond_dim = cond_dim
        in_out = list(zip(all_dims[:-1], all_dims[1:]))
        kw = dict(cond_dim=cond_dim, kernel_size=kernel_size, n_groups=n_groups, cond_predict_scale=cond_predict_scale)
        mid = all_dims[-1]
        self.local_cond_encoder = None
        self.mid_modules = nn.ModuleList([ConditionalResidualBlock1D(mid, mid, **kw),
                                          ConditionalResidualBlock1D(mid, mid, **kw)])
        self.down_modules = nn.ModuleList()
        for ind, (di, do) in enumerate(in_out):
            last = ind >= len(in_out) - 1
            self.down_modules.append(nn.ModuleList([ConditionalResidualBlock1D(di, do, **kw),
                                                    ConditionalResidualBlock1D(do, do, **kw),
                                                    Downsample1d(do) if not last else nn.Identity()]))
        self.up_modules = nn.ModuleList()
        for ind, (di, do) in enumerate(reversed(in_out[1:])):
            last = ind >= len(in_out) - 1  # never true (conditional_unet1d.py:186): every up level upsamples
            self.up_modules.append(nn.ModuleList([ConditionalResidualBlock1D(do * 2, di, **kw),
                                                  ConditionalResidualBlock1D(di, di, **kw),
                                                  Upsample1d(di) if not last else nn.Identity()]))
        self.final_conv = nn.Sequential(Conv1dBlock(down_dims[0], down_dims[0], kernel_size=kernel_size),
                                        Conv1dCL(down_dims[0], input_dim, 1))

    def forward(self, sample, timestep, local_cond=None, global_cond=None, **kwargs):
        """sample (B, T, input_dim), timestep (B,) or scalar tensor, global_cond (B, G) -> (B, T, input_dim).
        The reference transposes to (B, C, T) and back (conditional_unet1d.py:241,296); channel-last needs neither."""
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        elif timestep.dim() == 0:
            timestep = timestep[None].to(sample.device)
        timestep = timestep.expand(sample.shape[0])
        enc = self.diffusion_step_encoder
        g = PF.linear(enc[0](timestep), enc[1].weight, enc[1].bias)
        g = PF.linear(UF.mish(g), enc[3].weight, enc[3].bias)
        if global_cond is not None:
            g = torch.cat([g, global_cond], dim=-1)
        pad = -(-g.shape[1] // 8) * 8 - g.shape[1]
        mc = UF.mish(F.pad(g, (0, pad)) if pad else g)  # Mish(0) = 0: padding commutes with the activation
        x = sample.contiguous()
        h = []
        for r1, r2, down in self.down_modules:
            x = r2(r1(x, mc), mc)
            h.append(x)
            x = down(x)
        for m in self.mid_modules:
            x = m(x, mc)
        for r1, r2, up in self.up_modules:
            x = up(r2(r1(torch.cat((x, h.pop()), dim=-1), mc), mc))
        return self.final_conv[1](self.final_conv[0](x))


# ---- normalizer.py -------------------------------------------------------------------------------
class LinearNormalizer(nn.Module):
    """x * scale + offset per field (`_normalize`, normalizer.py:283-297); parameters live in
    `params_dict.<field>.{scale,offset}` like the reference's ParameterDict tree."""

    def __init__(self):
        super().__init__()
        self.params_dict = nn.ParameterDict()

    @torch.no_grad()
    def fit(self, data: dict, mode="limits", output_max=1.0, output_min=-1.0, range_eps=1e-4, fit_offset=True):
        """`_fit` (normalizer.py:195-280), last_n_dims = 1."""
        for key, v in data.items():
            v = torch.as_tensor(v, dtype=torch.float32)
            v = v.reshape(-1, v.shape[-1])
            lo, hi, mean, std = v.min(0).values, v.max(0).values, v.mean(0), v.std(0)
            if mode == "limits":
                if fit_offset:
                    rng = hi - lo
                    ignore = rng < range_eps
                    rng[ignore] = output_max - output_min
                    scale = (output_max - output_min) / rng
                    offset = output_min - scale * lo
                    offset[ignore] = (output_max + output_min) / 2 - lo[ignore]
                else:
                    out_abs = min(abs(output_min), abs(output_max))
                    in_abs = torch.maximum(lo.abs(), hi.abs())
                    in_abs[in_abs < range_eps] = out_abs
                    scale, offset = out_abs / in_abs, torch.zeros_like(mean)
            elif mode == "gaussian":
                scale = std.clone()
                scale[std < range_eps] = 1
                scale = 1 / scale
                offset = -mean * scale if fit_offset else torch.zeros_like(mean)
            else:
                raise ValueError(mode)
            stats = nn.ParameterDict({k: nn.Parameter(t, requires_grad=False)
                                      for k, t in (("min", lo), ("max", hi), ("mean", mean), ("std", std))})
            self.params_dict[key] = nn.ParameterDict({"scale": nn.Parameter(scale, requires_grad=False),
                                                      "offset": nn.Parameter(offset, requires_grad=False),
                                                      "input_stats": stats})
        return self

    def set_identity(self, dims: dict):
        for key, d in dims.items():
            self.params_dict[key] = nn.ParameterDict({"scale": nn.Parameter(torch.ones(d), requires_grad=False),
                                                      "offset": nn.Parameter(torch.zeros(d), requires_grad=False)})
        return self

    def normalize_field(self, key, x, forward=True):
        p = self.params_dict[key]
        if p["scale"].device != x.device:  # fields loaded / fit after the module was moved (reference: x goes to
            self.to(x.device)              # the parameters' device, normalizer.py:289; here the data's device wins)
            p = self.params_dict[key]
        scale, offset = p["scale"], p["offset"]
        shape = x.shape
        x = x.to(scale.dtype).reshape(-1, scale.shape[0])
        x = x * scale + offset if forward else (x - offset) / scale
        return x.reshape(shape)

    def normalize(self, x: dict):
        return {k: self.normalize_field(k, v) for k, v in x.items()}

    def _load_from_state_dict(self, state_dict, prefix, *args):  # dict_of_tensor_mixin.py:15-47
        root = prefix + "params_dict."
        dev = next((p.device for p in self.parameters()), None)
        for k, v in state_dict.items():
            if not k.startswith(root):
                continue
            keys = k[len(root):].split(".")
            dest = self.params_dict
            for name in keys[:-1]:
                if name not in dest:
                    dest[name] = nn.ParameterDict()
                dest = dest[name]
            dest[keys[-1]] = nn.Parameter(v.clone().to(dev) if dev is not None else v.clone(), requires_grad=False)


# ---- diffusers DDPMScheduler: forward process ------------------------------------------------------
class _SchedulerConfig:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class DDPMScheduler:
    """Training-time subset of `diffusers.schedulers.scheduling_ddpm.DDPMScheduler` (0.29.0) with the
    constructor arguments of maniskill2_diffusion_policy_model.yaml:29-38: beta schedule and `add_noise`."""

    def __init__(self, num_train_timesteps=100, beta_start=0.0001, beta_end=0.02, beta_schedule="squaredcos_cap_v2",
                 clip_sample=True, prediction_type="epsilon", variance_type="fixed_small", **_):
        if beta_schedule == "squaredcos_cap_v2":
            ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
            betas = torch.tensor([min(1 - ab((i + 1) / num_train_timesteps) / ab(i / num_train_timesteps), 0.999)
                                  for i in range(num_train_timesteps)], dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        if variance_type != "fixed_small":
            raise NotImplementedError(variance_type)
        self.config = _SchedulerConfig(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                                       clip_sample=clip_sample, clip_sample_range=1.0, variance_type=variance_type)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self._dev = {}

    def _coef(self, device):
        c = self._dev.get(device)
        if c is None:
            acp = self.alphas_cumprod.to(device)
            c = self._dev[device] = (acp ** 0.5, (1 - acp) ** 0.5)
        return c

    def set_timesteps(self, num_inference_steps):
        """"leading" spacing (the diffusers default): t_i = i * (T // n), descending."""
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        self.timesteps = torch.arange(self.num_inference_steps - 1, -1, -1) * ratio

    def step_coefficients(self, t: int):
        """The five scalars of one ancestral sampling step (Ho et al. 2020 eq. 6-7, 11; diffusers `step` with
        `variance_type="fixed_small"`), so that
            x0   = (x_t - c0 * eps_hat) * c1            [clamped to +-clip_sample_range when clip_sample]
            x_t' = c2 * x0 + c3 * x_t + c4 * noise      [c4 = 0 at t = 0]
        Computed in fp32 from `alphas_cumprod` exactly as the scheduler does."""
        n = getattr(self, "num_inference_steps", None) or self.config.num_train_timesteps
        prev_t = t - self.config.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        var = torch.clamp(b_prev / b_t * cur_b, min=1e-20)
        return torch.stack([b_t ** 0.5, 1.0 / a_t ** 0.5, (a_prev ** 0.5 * cur_b) / b_t, cur_a ** 0.5 * b_prev / b_t,
                            var ** 0.5 if t > 0 else torch.tensor(0.0)]).float()

    def add_noise(self, x, noise, timesteps):
        a, s = self._coef(x.device)
        a, s = a[timesteps], s[timesteps]
        while a.dim() < x.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * x + s * noise


# ---- pcd_obs_encoder.py --------------------------------------------------------------------------
class PCDObsEncoder(nn.Module):
    def __init__(self, shape_meta: dict, pcd_model, share_pcd_model: bool = True, n_obs_step: int = 2,
                 pcd_nsample: int = 16, pcd_npoints: int = 1024, use_mask: bool = False, bg_ratio: float = 0.0,
                 pcd_hidden_dim: int = 128, projector_layers: int = 2, projector_channels=(128, 128, 128),
                 pre_sample=False, in_channel=6, **kwargs):
        super().__init__()
        if not share_pcd_model or not isinstance(pcd_model, nn.Module):
            raise NotImplementedError("per-key point-cloud models (share_pcd_model=False) are not used by any config")
        self._dummy_variable = nn.Parameter(torch.empty(0))  # module_attr_mixin.py:7-9 (state_dict key)
        self.key_model_map = nn.ModuleDict({"pcd": pcd_model})
        obs = shape_meta["obs"]
        self.pcd_keys = sorted(k for k, a in obs.items() if a.get("type", "low_dim") == "pcd")
        self.low_dim_keys = sorted(k for k, a in obs.items() if a.get("type", "low_dim") == "low_dim")
        for k, a in obs.items():
            if a.get("type", "low_dim") not in ("pcd", "low_dim"):
                raise RuntimeError(f"Unsupported obs type: {a.get('type')}")
        self.key_shape_map = {k: tuple(a["shape"]) for k, a in obs.items()}
        self.shape_meta, self.share_pcd_model, self.n_obs_step = shape_meta, share_pcd_model, n_obs_step
        self.pcd_nsample, self.pcd_npoints, self.use_mask, self.bg_ratio, self.pre_sample = (
            pcd_nsample, pcd_npoints, use_mask, bg_ratio, pre_sample)
        if not pre_sample:
            self.linear = nn.Linear(3 + pcd_model.num_channels, pcd_hidden_dim, bias=False)
            self.bn = nn.BatchNorm1d(pcd_hidden_dim)
        else:  # pcd_obs_encoder.py:91-93: the head runs on the raw channels, in front of the backbone
            self.linear = nn.Linear(3 + in_channel, in_channel, bias=False)
            self.bn = nn.BatchNorm1d(in_channel)
        self.pool = nn.MaxPool1d(pcd_nsample)
        self.relu = nn.ReLU(inplace=True)
        proj = []
        for i in range(projector_layers):
            cin = pcd_model.num_channels if (i == 0 and pre_sample) else pcd_hidden_dim  # :101-110
            proj += [nn.Conv1d(cin, projector_channels[i], kernel_size=1), nn.BatchNorm1d(projector_channels[i]),
                     nn.ReLU(inplace=True)]
        proj += [nn.MaxPool1d(pcd_npoints), nn.Conv1d(projector_channels[i], projector_channels[i + 1], kernel_size=1),
                 nn.BatchNorm1d(projector_channels[i + 1])]
        self.projector = nn.Sequential(*proj)  # parameter containers; forward below runs token-major
        self.projector_layers, self.projector_channels = projector_layers, list(projector_channels)

    def output_shape(self):
        return (self.projector_channels[-1] + sum(int(self.key_shape_map[k][0]) for k in self.low_dim_keys),)

    def pcd_sampling(self, pxo, mask, hints):
        """pcd_obs_encoder.py:123-198 -> (new coords, pooled features (b*M, c), new offsets, picked indices)."""
        from .act import sample_indices

        p, x, o = pxo
        b = o.shape[0]
        n_o = torch.arange(1, b + 1, dtype=torch.int32, device=o.device) * self.pcd_npoints
        o32 = o.int() if o.dtype != torch.int32 else o
        idx = sample_indices(p, o32, self.pcd_npoints, mask if self.use_mask else None, self.bg_ratio, hints)
        n_p = p[idx.long(), :].contiguous()
        knn_idx, _ = pointops.ops.KNNQuery.apply(self.pcd_nsample, p, o32, n_p, n_o, False)
        return n_p, PF.set_abstraction(p, x, o32, n_p, n_o, knn_idx, self.linear.weight, self.bn), n_o, idx

    def encode_pcd(self, pcd_model, pcd):
        b = pcd["offset"].shape[0]
        mask = pcd.get("mask", None) if self.use_mask else None
        hints = {k: pcd.get(k, None) for k in ("n_max", "fg_n_max", "bg_n_max")}
        if self.pre_sample:  # pcd_obs_encoder.py:201-218 (works on a copy: the caller's dict is not rewritten)
            coord, feats, offset, idx = self.pcd_sampling((pcd["coord"], pcd["feat"], pcd["offset"]), mask, hints)
            x = pcd_model(dict(pcd, coord=coord, feat=feats, offset=offset, grid_coord=pcd["grid_coord"][idx.long()]))
        else:
            feats = pcd_model(pcd)
            _, x, _, _ = self.pcd_sampling((pcd["coord"], feats, pcd["offset"]), mask, hints)
        # projector (pcd_obs_encoder.py:100-121): 1x1 Conv1d + BN + ReLU per point, max over the M points,
        # 1x1 Conv1d + BN -- on token-major rows (a 1x1 convolution is a Linear)
        for i in range(self.projector_layers):
            conv, bn = self.projector[3 * i], self.projector[3 * i + 1]
            if bn.training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            x = PF.batchnorm_relu(PF.linear(x, conv.weight.squeeze(-1), conv.bias), bn)
        x = x.view(b, self.pcd_npoints, -1).amax(dim=1)
        conv, bn = self.projector[3 * self.projector_layers + 1], self.projector[3 * self.projector_layers + 2]
        x = PF.linear(x, conv.weight.squeeze(-1), conv.bias)
        if bn.training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training,
                            bn.momentum if bn.momentum is not None else 0.0, bn.eps)

    def forward(self, obs_dict):
        batch_size = None
        feats = []
        for key in self.pcd_keys:
            pcd = obs_dict[key]
            batch_size = len(pcd["offset"])
            assert batch_size % self.n_obs_step == 0, (batch_size, self.n_obs_step)
            assert tuple(pcd["feat"].shape[1:]) == self.key_shape_map[key]
            feats.append(self.encode_pcd(self.key_model_map["pcd"], pcd).reshape(batch_size, -1))
        for key in self.low_dim_keys:
            data = obs_dict[key]
            assert batch_size is None or data.shape[0] == batch_size, (key, batch_size, data.shape)
            batch_size = data.shape[0]
            feats.append(data)
        return torch.cat(feats, dim=-1)


# ---- diffusion_unet_image_policy.py --------------------------------------------------------------
class _MaskGeneratorStub(nn.Module):
    """`LowdimMaskGenerator(obs_dim=0, action_visible=False)` (mask_generator.py:41-105): with observations as
    global conditioning the in-painting mask is identically False -- nothing to generate; only its
    `_dummy_variable` state_dict key remains."""

    def __init__(self):
        super().__init__()
        self._dummy_variable = nn.Parameter(torch.empty(0))


class DiffusionUnetImagePolicy(nn.Module):
    def __init__(self, shape_meta: dict, noise_scheduler, obs_encoder, horizon, n_action_steps, n_obs_steps,
                 num_inference_steps=None, obs_as_global_cond=True, diffusion_step_embed_dim=256,
                 down_dims=(256, 512, 1024), kernel_size=5, n_groups=8, cond_predict_scale=True, **kwargs):
        super().__init__()
        if not obs_as_global_cond:
            raise NotImplementedError  # as the reference's compute_loss (diffusion_unet_image_policy.py:259-260)
        self._dummy_variable = nn.Parameter(torch.empty(0))
        action_dim = int(shape_meta["action"]["shape"][0])
        obs_feature_dim = obs_encoder.output_shape()[0]
        global_cond_dim = obs_feature_dim * n_obs_steps
        if shape_meta.get("goal") is not None:
            global_cond_dim += int(shape_meta["goal"]["task_emb"]["shape"][0])
        self.obs_encoder = obs_encoder
        self.model = ConditionalUnet1D(input_dim=action_dim, local_cond_dim=None, global_cond_dim=global_cond_dim,
                                       diffusion_step_embed_dim=diffusion_step_embed_dim, down_dims=down_dims,
                                       kernel_size=kernel_size, n_groups=n_groups, cond_predict_scale=cond_predict_scale)
        self.noise_scheduler = noise_scheduler
        self.mask_generator = _MaskGeneratorStub()
        self.normalizer = LinearNormalizer()
        self.horizon, self.obs_feature_dim, self.action_dim = horizon, obs_feature_dim, action_dim
        self.n_action_steps, self.n_obs_steps, self.obs_as_global_cond = n_action_steps, n_obs_steps, obs_as_global_cond
        self.num_inference_steps = num_inference_steps or noise_scheduler.config.num_train_timesteps
        self.kwargs = kwargs
        self._sample_graphs = {}  # (trajectory shape, cond shape, device) -> (CUDAGraph, static tensors)
        self._sample_graphs_fp = None  # weight fingerprint the cached graphs were captured under

    def set_normalizer(self, normalizer):
        self.normalizer.load_state_dict(normalizer.state_dict())

    def compute_loss(self, batch):
        """diffusion_unet_image_policy.py:233-313.  Test hooks: `batch["_noise"]`, `batch["_timesteps"]` replace
        the two random draws (:281-290).  The batch is not mutated (the reference pops `obs.pcds`)."""
        assert "valid_mask" not in batch
        obs = {k: v for k, v in batch["obs"].items() if k != "pcds"}
        pcds = batch["obs"].get("pcds", None)
        nobs = self.normalizer.normalize(obs)
        nactions = self.normalizer.normalize_field("action", batch["action"])
        bs = nactions.shape[0]
        this_nobs = {k: v[:, : self.n_obs_steps].reshape(-1, *v.shape[2:]) for k, v in nobs.items()}
        if pcds is not None:
            this_nobs["pcds"] = pcds
        global_cond = self.obs_encoder(this_nobs).reshape(bs, -1)
        goal = batch.get("goal", None)
        if goal is not None:
            if "task_emb" not in goal:
                raise NotImplementedError("image goals belong to the image policy, not the point-cloud path")
            global_cond = torch.cat([global_cond, goal["task_emb"]], dim=-1)
        noise = batch["_noise"] if "_noise" in batch else torch.randn(nactions.shape, device=nactions.device)
        if "_timesteps" in batch:
            timesteps = batch["_timesteps"]
        else:
            timesteps = torch.randint(0, self.noise_scheduler.config.num_train_timesteps, (bs,), device=nactions.device).long()
        noisy = self.noise_scheduler.add_noise(nactions, noise, timesteps)
        PF.grad_boundary(global_cond, "model")  # d(global_cond) available = the denoiser's (255 M) gradients are final
        pred = self.model(noisy, timesteps, local_cond=None, global_cond=global_cond)
        pred_type = self.noise_scheduler.config.prediction_type
        if pred_type == "epsilon":
            target = noise
        elif pred_type == "sample":
            target = nactions
        else:
            raise ValueError(f"Unsupported prediction type {pred_type}")
        loss = F.mse_loss(pred, target, reduction="none").reshape(bs, -1).mean(dim=1).mean()
        return dict(loss=loss)

    def grad_buckets(self):
        """See act.ACTPCD.grad_buckets: the denoiser (the ~1 GB gradient) finishes its backward before the encoder starts."""
        return [("model", ("model.",)), (None, ("",))]

    def sync_free(self, pcds) -> bool:
        """True when the host-known cloud-size hints let FPS run without a device->host read."""
        enc = self.obs_encoder
        need = ["n_max"] if not (enc.use_mask and pcds.get("mask", None) is not None) else (
            ["fg_n_max"] + (["bg_n_max"] if enc.bg_ratio > 0.0 else []))
        return all(pcds.get(k, None) is not None for k in need)

    # ========= inference (diffusion_unet_image_policy.py:106-231) =========
    def _denoise_step(self, traj, t, coef, noise, global_cond):
        """One reverse-diffusion step: denoiser forward + the scheduler update, all on device tensors (timestep,
        the 5 step coefficients and the noise are INPUTS), so the whole step can be captured once and replayed."""
        out = self.model(traj, t, local_cond=None, global_cond=global_cond)
        if self.noise_scheduler.config.prediction_type == "epsilon":
            x0 = (traj - coef[0] * out) * coef[1]
        else:
            x0 = out
        if self.noise_scheduler.config.clip_sample:
            r = self.noise_scheduler.config.clip_sample_range
            x0 = x0.clamp(-r, r)
        return coef[2] * x0 + coef[3] * traj + coef[4] * noise

    @torch.no_grad()
    def conditional_sample(self, shape, global_cond, noises=None, use_cuda_graph=True):
        """The reference's sampling loop (:106-146) with observations as global conditioning (its in-painting mask
        is identically False, so the two masked assignments are no-ops).  `noises` (iterable: x_T, then one
        tensor per step with t > 0) replaces the sampler's draws -- test hook.
        B200 path: ONE denoising step (~190 kernels) is captured into a CUDA graph the first time a shape is seen
        and replayed `num_inference_steps` times; per step only the timestep, 5 coefficients and the noise are
        copied into the graph's static inputs."""
        sch = self.noise_scheduler
        sch.set_timesteps(self.num_inference_steps)
        dev = global_cond.device
        it = iter(noises) if noises is not None else None
        traj = (next(it).to(dev) if it is not None else torch.randn(shape, device=dev)).contiguous()
        ts = sch.timesteps.to(dev)
        coefs = torch.stack([sch.step_coefficients(int(t)) for t in sch.timesteps]).to(dev)
        key = (tuple(shape), tuple(global_cond.shape), str(dev))
        if use_cuda_graph:
            # A captured step bakes in ADDRESSES: the fp32 biases / GroupNorm parameters and the bf16 operand copies
            # `functional._wb` made of the weights at capture time.  Any change of the weights behind those addresses
            # (load_state_dict, an in-place torch.optim step: both bump Parameter._version; .to() / re-assignment:
            # new data_ptr) would leave the graph computing with the weights of its first call, so the cache is
            # keyed on a fingerprint of every denoiser parameter and dropped when it changes.  Under a BCTrainer the
            # operands are views of the flat buffers the fused AdamW kernel updates in place (same address, current
            # values, no version bump) -- the graph stays valid there by construction.
            fp = tuple((p_.data_ptr(), p_._version) for p_ in self.model.parameters())
            if fp != self._sample_graphs_fp:
                self._sample_graphs.clear()
                self._sample_graphs_fp = fp
        entry = self._sample_graphs.get(key) if use_cuda_graph else None
        if use_cuda_graph and entry is None:
            st = dict(traj=traj.clone(), t=ts[0].clone(), coef=coefs[0].clone(), noise=torch.zeros(shape, device=dev),
                      cond=global_cond.clone())
            side = torch.cuda.Stream()  # warm-up off the capture: lazy initialisations, allocator growth
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._denoise_step(st["traj"], st["t"], st["coef"], st["noise"], st["cond"])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["out"] = self._denoise_step(st["traj"], st["t"], st["coef"], st["noise"], st["cond"])
            entry = self._sample_graphs[key] = (graph, st)
        for i, t in enumerate(sch.timesteps.tolist()):
            noise = None
            if t > 0:
                noise = next(it).to(dev) if it is not None else torch.randn(shape, device=dev)
            if entry is not None:
                graph, st = entry
                if i == 0:
                    st["cond"].copy_(global_cond)
                st["traj"].copy_(traj)
                st["t"].copy_(ts[i])
                st["coef"].copy_(coefs[i])
                if noise is not None:
                    st["noise"].copy_(noise)
                graph.replay()
                traj = st["out"]
            else:
                traj = self._denoise_step(traj, ts[i], coefs[i], noise if noise is not None else torch.zeros_like(traj),
                                          global_cond)
        return traj.clone()

    @torch.no_grad()
    def predict_action(self, obs_dict, noises=None, use_cuda_graph=True):
        """obs_dict = {"obs": {qpos (B, To.., Q), pcds: {...}}[, "goal": {task_emb}]} -> {"action", "action_pred"}
        (:148-231).  The input dict is not mutated (the reference pops `pcds`)."""
        assert "past_action" not in obs_dict
        src = obs_dict["obs"] if "obs" in obs_dict else obs_dict
        pcds = src.get("pcds", None)
        nobs = self.normalizer.normalize({k: v for k, v in src.items() if k != "pcds"})
        B = next(iter(nobs.values())).shape[0]
        this_nobs = {k: v[:, : self.n_obs_steps].reshape(-1, *v.shape[2:]) for k, v in nobs.items()}
        if pcds is not None:
            this_nobs["pcds"] = pcds
        global_cond = self.obs_encoder(this_nobs).reshape(B, -1)
        goal = obs_dict.get("goal", None)
        if goal is not None:
            if "task_emb" not in goal:
                raise NotImplementedError("image goals belong to the image policy, not the point-cloud path")
            global_cond = torch.cat([global_cond, goal["task_emb"]], dim=-1)
        nsample = self.conditional_sample((B, self.horizon, self.action_dim), global_cond.contiguous(), noises, use_cuda_graph)
        action_pred = self.normalizer.normalize_field("action", nsample[..., : self.action_dim], forward=False)
        start = self.n_obs_steps - 1
        return {"action": action_pred[:, start:start + self.n_action_steps], "action_pred": action_pred}

    def forward(self, batch):
        """maniskill2_dp_bc_module.py:59-63: training -> compute_loss, evaluation -> predict_action."""
        return self.compute_loss(batch) if self.training else self.predict_action(batch)


def build_dp_policy(cfg: dict):
    """Convenience constructor with the structure of exp_maniskill2_diffusion_policy/.../scratch_pointnet_pcd.yaml +
    maniskill2_diffusion_policy_model.yaml."""
    from .pointnet import PointNet

    shape_meta = {"obs": {"pcds": {"shape": [6], "type": "pcd"}, "qpos": {"shape": [cfg["qpos_dim"]], "type": "low_dim"}},
                  "action": {"shape": [cfg["action_dim"]]}, "goal": None}
    if cfg.get("goal_dim", 0):
        shape_meta["goal"] = {"task_emb": {"shape": [cfg["goal_dim"]]}}
    enc = PCDObsEncoder(shape_meta, PointNet(6, cfg["backbone_classes"]), share_pcd_model=True, n_obs_step=cfg["n_obs_steps"],
                        pcd_nsample=cfg["pcd_nsample"], pcd_npoints=cfg["pcd_npoints"], pcd_hidden_dim=cfg["pcd_hidden_dim"],
                        projector_layers=cfg["projector_layers"], projector_channels=cfg["projector_channels"])
    return DiffusionUnetImagePolicy(shape_meta, DDPMScheduler(num_train_timesteps=cfg.get("num_train_timesteps", 100)), enc,
                                    horizon=cfg["horizon"], n_action_steps=cfg.get("n_action_steps", 8),
                                    n_obs_steps=cfg["n_obs_steps"], num_inference_steps=cfg.get("num_inference_steps", None),
                                    diffusion_step_embed_dim=cfg["diffusion_step_embed_dim"],
                                    down_dims=cfg["down_dims"], kernel_size=cfg["kernel_size"], n_groups=cfg["n_groups"],
                                    cond_predict_scale=cfg.get("cond_predict_scale", True))


# BASELINE.json configs 3 / 5 with the PointNet backbone (scratch_pointnet_pcd.yaml; SpUNet is SURVEY.md 8f item 1)
DP_MODEL_CFG = dict(qpos_dim=9, action_dim=7, backbone_classes=96, n_obs_steps=2, pcd_nsample=16, pcd_hidden_dim=96,
                    projector_layers=1, projector_channels=[96, 128, 128], horizon=16, diffusion_step_embed_dim=128,
                    down_dims=[512, 1024, 2048], kernel_size=5, n_groups=8, cond_predict_scale=True, goal_dim=0)
