"""ORACLE (test infrastructure, not product code).

CPU fp32 restatement of the reference's reverse-time bridge sampling loop for ONE image, citing
/root/reference/text-guided/inversion/{p2p_h_edit,inversion_utils,ddpm_inversion}.py.  It is the checker for the
CUDA path (tests/) and the timed CPU arm of bench.py (`cpu_baseline`, `--impl reference`); it never runs on the
product path.  Pinned against the reference's own functions by tests/test_oracle_pin.py and tests/golden/.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .p2p import EditSpec, P2PState, install_oracle_p2p, local_blend


# ---- scheduler algebra ------------------------------------------------------------------------------------
def _abar_prev(sched, t: int):
    """inversion_utils.py:84-87 -- p = t - 1000//T ; final_alpha_cumprod when p < 0."""
    p = int(t) - sched.config.num_train_timesteps // sched.num_inference_steps
    return sched.alphas_cumprod[p] if p >= 0 else sched.final_alpha_cumprod


def variance(sched, t: int):
    """inversion_utils.py:38-56."""
    a_t, a_p = sched.alphas_cumprod[int(t)], _abar_prev(sched, t)
    return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)


def reverse_step(sched, eps, t: int, x, eta, z, is_ddim_inversion: bool):
    """inversion_utils.py:58-126 (variance_noise always supplied on the h-Edit paths)."""
    a_t, a_p = sched.alphas_cumprod[int(t)], _abar_prev(sched, t)
    x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    var = variance(sched, t)
    if is_ddim_inversion:
        direction = (1 - a_p) ** 0.5 * eps
    else:
        direction = (1 - a_p - (eta ** 2) * var) ** 0.5 * eps
    prev = a_p ** 0.5 * x0 + direction
    if eta > 0:
        prev = prev + (eta * z if is_ddim_inversion else eta * var ** 0.5 * z)
    return prev


def full_coeff(sched, t: int, tt: int, eta, is_ddim_inversion: bool):
    """inversion_utils.py:168-195 -- sqrt(1 - abar_tt - omega^2)."""
    ab = sched.alphas_cumprod
    sig, a = (1 - ab) ** 0.5, ab ** 0.5
    omega = eta * (sig[int(tt)] / (sig[int(t)] * a[int(tt)])) * ((ab[int(tt)] - ab[int(t)]) ** 0.5)
    if is_ddim_inversion:
        omega = 0
    return (1 - ab[int(tt)] - omega ** 2) ** 0.5


# ---- inversion --------------------------------------------------------------------------------------------
@torch.no_grad()
def ddpm_inversion(unet, sched, ctx_uncond, ctx_src, x0, eta: float, cfg_src: float, generator: torch.Generator):
    """ddpm_inversion.py:5-52 (independent x_t draws) + :54-167 (solve z_t).  Noise comes from `generator`
    (the reference is unseeded, :20-21)."""
    T = sched.num_inference_steps
    ab = sched.alphas_cumprod
    ts = sched.timesteps
    pos = {int(v): k for k, v in enumerate(ts)}
    xts = torch.zeros((T + 1,) + tuple(x0.shape[1:]))
    xts[0] = x0[0]
    for t in reversed(ts):
        idx = T - pos[int(t)]
        noise = torch.randn(x0.shape, generator=generator)
        xts[idx] = (x0 * ab[t] ** 0.5 + noise * (1 - ab[t]) ** 0.5)[0]
    zs = torch.zeros((T,) + tuple(x0.shape[1:]))
    for t in ts:
        idx = T - pos[int(t)] - 1
        xt = xts[idx + 1][None]
        e_u = unet(xt, t, encoder_hidden_states=ctx_uncond).sample
        e_c = unet(xt, t, encoder_hidden_states=ctx_src).sample
        eps = e_u + cfg_src * (e_c - e_u)
        x0_hat = (xt - (1 - ab[t]) ** 0.5 * eps) / ab[t] ** 0.5
        a_p = _abar_prev(sched, int(t))
        var = variance(sched, int(t))
        mu = a_p ** 0.5 * x0_hat + (1 - a_p - (eta ** 2) * var) ** 0.5 * eps
        z = (xts[idx][None] - mu) / (eta * var ** 0.5)
        zs[idx] = z[0]
        xts[idx] = (mu + (eta * var ** 0.5) * z)[0]
    return zs, xts


# ---- the north-star loop ----------------------------------------------------------------------------------
@torch.no_grad()
def h_edit_p2p_implicit(unet, sched, ctx_uncond, ctx_src, ctx_tar, xT, zs, spec: Optional[EditSpec],
                        cfg_scales: Sequence[float], eta: float = 1.0, weight_reconstruction: float = 0.075,
                        optimization_steps: int = 1, after_skip_steps: Optional[int] = None,
                        is_ddim_inversion: bool = False, trace: Optional[List[torch.Tensor]] = None):
    """p2p_h_edit.py:529-701.  ctx_* are (1,77,D).  Returns (edited, reconstructed) each (1,C,H,W).
    `trace`, when given, receives xt (2,C,H,W) after every timestep (post LocalBlend)."""
    T = sched.num_inference_steps
    S = T if after_skip_steps is None else after_skip_steps
    w_src, w_src_edit, w_tar = [float(c) for c in cfg_scales]
    ab = sched.alphas_cumprod
    op = [int(t) for t in sched.timesteps[-S:]]
    pos = {t: k for k, t in enumerate(op)}
    state = P2PState()
    if spec is not None:
        install_oracle_p2p(unet, state, spec)
    off = {"use_controller": False}
    xt = torch.cat([xT.reshape(1, *xT.shape[-3:])] * 2)

    for i, t in enumerate(op):
        idx = T - pos[t] - (T - S + 1)
        z = zs[idx]
        # call A (:606-616): [xo, xe] x [null, src], P2P off
        out = unet(torch.cat([xt, xt]), t, encoder_hidden_states=torch.cat([ctx_uncond, ctx_uncond, ctx_src, ctx_src]),
                   cross_attention_kwargs=off).sample
        e_u, e_c = out.chunk(2)
        prev = reverse_step(sched, e_u + w_src * (e_c - e_u), t, xt, eta, z, is_ddim_inversion)     # :619
        x_orig, x_base = prev.chunk(2)
        tt = op[i + 1] if i < len(op) - 1 else 0                                                        # :626-629
        x_opt = x_base.clone()
        for k in range(optimization_steps):
            save = not (k < optimization_steps - 1 and optimization_steps > 1)                          # :637-640
            c_src = unet(x_opt, tt, encoder_hidden_states=ctx_src, cross_attention_kwargs=off).sample   # call B :644
            out = unet(torch.cat([x_orig, x_opt, x_orig, x_opt]), tt,                                    # call C :652
                       encoder_hidden_states=torch.cat([ctx_uncond, ctx_uncond, ctx_src, ctx_tar]),
                       cross_attention_kwargs={"save_attn": save} if spec is not None else off).sample
            u_tar, c_tar = out[1:2], out[3:4]
            eps_src_edit = u_tar + w_src_edit * (c_src - u_tar)                                          # :659
            eps_tar = u_tar + w_tar * (c_tar - u_tar)                                                    # :660
            coeff = full_coeff(sched, t, tt, eta, is_ddim_inversion) - (1 - ab[t]) ** 0.5 * (ab[tt] ** 0.5 / ab[t] ** 0.5)
            corr = eps_tar - eps_src_edit                                                                # :667
            if k > 0:                                                                                    # :670-684
                g = torch.sign(x_opt - x_base) / x_opt.numel()       # d/dx mean|x - x_base|
                c_n = float((corr * corr).mean().sqrt())
                g_n = float((g * g).mean().sqrt())
                rec = x_opt - (c_n / (g_n + 1e-8) * weight_reconstruction) * g
            else:
                rec = x_opt
            x_opt = rec + coeff * corr                                                                   # :689-692
        xt = torch.cat([x_orig, x_opt])
        if spec is not None:
            xt = local_blend(state, spec, xt)                                                            # :698-699
        if trace is not None:
            trace.append(xt.clone())
    return xt[1:2].clone(), xt[0:1].clone()
