"""Inversion callables with the reference's names and return conventions, running the denoiser on the native engine:
  inversion_forward_process_ddpm  <- text-guided/inversion/ddpm_inversion.py:54  (edit-friendly DDPM inversion)
  ddim_inversion                  <- text-guided/inversion/ddim_inversion.py:55   (deterministic inversion for h-Edit-D)

B200-first restructuring: given the independently sampled x_t's, the T steps of the DDPM inversion do not depend on each
other (the reference's in-place rewrite of xts[idx], ddpm_inversion.py:161-162, is a round-off-level re-derivation), so
all 2T noise predictions are ONE batched UNet launch with per-sample timesteps instead of 2T launches of batch 1.  The
second sweep of the DDIM inversion is batched the same way; its first sweep is inherently sequential.
The scalar/elementwise algebra is a handful of fused torch ops on the device (host orchestration)."""
from __future__ import annotations

import torch

from .samplers import encode_text, get_engine


def _prev_alpha(sched, t: int):
    p = int(t) - sched.config.num_train_timesteps // sched.num_inference_steps
    return sched.alphas_cumprod[p] if p >= 0 else sched.final_alpha_cumprod


def _variance(sched, t: int):
    a_t, a_p = sched.alphas_cumprod[int(t)], _prev_alpha(sched, t)
    return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)


def sample_xts_from_x0(model, x0, num_inference_steps=50):
    """ddpm_inversion.py:5-52: independent draws x_t ~ q(x_t | x_0) (global RNG, like the reference)."""
    ab = model.scheduler.alphas_cumprod
    ts = model.scheduler.timesteps
    pos = {int(v): k for k, v in enumerate(ts)}
    shape = (num_inference_steps + 1,) + tuple(x0.shape[1:])
    xts = torch.zeros(shape, device=x0.device)
    noise_added = torch.zeros(shape, device=x0.device)
    xts[0] = x0[0]
    for t in reversed(ts):
        idx = num_inference_steps - pos[int(t)]
        noise = torch.randn_like(x0)
        xts[idx] = (x0 * (ab[int(t)] ** 0.5) + noise * ((1 - ab[int(t)]) ** 0.5))[0]
        noise_added[idx] = noise[0]
    return xts, noise_added


@torch.no_grad()
def inversion_forward_process_ddpm(model, x0, etas=None, prog_bar=True, prompt="", cfg_scale_src=1.0, cfg_scale_src_edit=3.5,
                                   num_inference_steps=50):
    """Reference signature; returns (xt, zs, xts, noise_added)."""
    assert not (etas is None or (isinstance(etas, (int, float)) and etas == 0)), "eta must be > 0 (reference assert, ddpm_inversion.py:139)"
    T = model.scheduler.num_inference_steps
    etas = [etas] * T if isinstance(etas, (int, float)) else list(etas)
    sched = model.scheduler
    ts = [int(t) for t in sched.timesteps]
    xts, noise_added = sample_xts_from_x0(model, x0, num_inference_steps=num_inference_steps)
    eng = get_engine(model, max_samples=max(5, min(2 * T, 40)))
    dev = torch.device("cuda", eng.device)
    ctx = [encode_text(model, "")]
    if prompt != "":
        ctx.append(encode_text(model, prompt))
    ctx = torch.cat(ctx).float().to(dev)
    # step k (timestep ts[k]) starts from xts[T - k]
    x_in = torch.stack([xts[T - k] for k in range(T)]).to(dev)
    n_ctx = ctx.shape[0]
    x_all = torch.cat([x_in] * n_ctx)
    t_all = ts * n_ctx
    c_idx = [c for c in range(n_ctx) for _ in range(T)]
    eps = eng.forward(x_all, t_all, ctx, ctx_index=c_idx)
    e_u = eps[:T]
    noise_pred = e_u + cfg_scale_src * (eps[T:] - e_u) if n_ctx == 2 else e_u
    ab = sched.alphas_cumprod
    zs = torch.zeros((T,) + tuple(x0.shape[1:]), device=x0.device)
    for k, t in enumerate(ts):
        idx = T - k - 1
        xt = x_in[k]
        a_t, a_p, var = ab[t], _prev_alpha(sched, t), _variance(sched, t)
        x0_hat = (xt - (1 - a_t) ** 0.5 * noise_pred[k]) / a_t ** 0.5
        mu = a_p ** 0.5 * x0_hat + (1 - a_p - (etas[idx] ** 2) * var) ** 0.5 * noise_pred[k]
        sig = etas[idx] * var ** 0.5
        z = (xts[idx].to(dev) - mu) / sig
        zs[idx] = z.to(x0.device)
        xts[idx] = (mu + sig * z).to(x0.device)
    return xts[1][None], zs, xts, noise_added


@torch.no_grad()
def ddim_inversion(model, w0, prompt: str, cfg_scale: float):
    """Reference signature; returns (latent, zs, latents)."""
    sched = model.scheduler
    T = sched.num_inference_steps
    ts = [int(t) for t in sched.timesteps]
    ratio = sched.config.num_train_timesteps // T
    ab = sched.alphas_cumprod
    eng = get_engine(model, max_samples=max(5, min(2 * T, 40)))
    dev = torch.device("cuda", eng.device)
    ctx = torch.cat([encode_text(model, ""), encode_text(model, prompt)]).float().to(dev)

    def noise_pred(x, tlist):          # x (n,C,h,w) at timesteps tlist -> CFG noise prediction
        n = x.shape[0]
        eps = eng.forward(torch.cat([x, x]), list(tlist) * 2, ctx, ctx_index=[0] * n + [1] * n)
        return eps[:n] + cfg_scale * (eps[n:] - eps[:n])

    latent = w0.clone().detach().to(dev)
    latents = [latent]
    for i in range(T):                 # sequential sweep x_0 -> x_T (ddim_inversion.py:8-29,83-87)
        t = ts[T - i - 1]
        e = noise_pred(latent, [t])
        tp = min(t - ratio, 999)
        a_t = ab[tp] if tp >= 0 else sched.final_alpha_cumprod
        a_n = ab[t]
        x0_hat = (latent - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
        latent = a_n ** 0.5 * x0_hat + (1 - a_n) ** 0.5 * e
        latents.append(latent)
    # second sweep (ddim_inversion.py:98-127): z_t = x_{t-1} - mu(x_t); every step only reads `latents` -> one batched launch
    x_in = torch.cat([latents[T - k] for k in range(T)])
    e_all = noise_pred(x_in, ts)
    zs = torch.zeros((T,) + tuple(w0.shape[1:]), device=w0.device)
    for k, t in enumerate(ts):
        idx = T - k - 1
        xt = x_in[k:k + 1]
        a_p = _prev_alpha(sched, t)
        x0_hat = (xt - (1 - ab[t]) ** 0.5 * e_all[k:k + 1]) / ab[t] ** 0.5
        mu = a_p ** 0.5 * x0_hat + (1 - a_p) ** 0.5 * e_all[k:k + 1]
        z = latents[idx] - mu
        zs[idx] = z[0].to(w0.device)
        latents[idx] = mu + z
    latents = [x.to(w0.device) for x in latents]
    return latents[-1], zs, latents
