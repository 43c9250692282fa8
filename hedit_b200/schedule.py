"""Per-step scalar tables of the h-Edit update, computed on the host ONCE per edit in the reference's own fp32
operation order (text-guided/inversion/inversion_utils.py:38-56 get_variance, :58-126 reverse_step, :168-195
compute_full_coeff; coefficient line text-guided/inversion/p2p_h_edit.py:664-665).  The reference re-derives these
every step with CPU-tensor indexing by CUDA scalars (implicit host syncs); here they become a (steps,) table that the
fused element kernels read."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def step_tables(scheduler, after_skip_steps: int, eta, is_ddim_inversion: bool) -> Tuple[List[int], np.ndarray]:
    """Returns (timesteps [steps+1] incl. the final previous timestep 0, coef [steps,6] float32)."""
    T = scheduler.num_inference_steps
    etas = [eta] * T if isinstance(eta, (int, float)) else list(eta)
    ac = scheduler.alphas_cumprod.detach().float().cpu()
    final = torch.as_tensor(scheduler.final_alpha_cumprod).float().cpu()
    ts = [int(t) for t in scheduler.timesteps[-after_skip_steps:]]
    ratio = scheduler.config.num_train_timesteps // T
    sig, a = (1 - ac) ** 0.5, ac ** 0.5
    coef = np.zeros((len(ts), 6), dtype=np.float32)
    for i, t in enumerate(ts):
        idx = T - i - (T - after_skip_steps + 1)          # p2p_h_edit.py:599 with t_to_idx[t] == i
        e = etas[idx]
        tt = ts[i + 1] if i < len(ts) - 1 else 0
        p = t - ratio
        a_t = ac[t]
        a_p = ac[p] if p >= 0 else final
        var = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        if is_ddim_inversion:
            direction = (1 - a_p) ** 0.5
            noise = torch.tensor(float(e))
            omega = 0
        else:
            direction = (1 - a_p - (e ** 2) * var) ** 0.5
            noise = e * var ** 0.5
            omega = e * (sig[tt] / (sig[t] * a[tt])) * ((ac[tt] - ac[t]) ** 0.5)
        if not e > 0:
            noise = torch.tensor(0.0)
        full = (1 - ac[tt] - omega ** 2) ** 0.5
        coeff = full - sig[t] * (a[tt] / a[t])
        coef[i] = [float((1 - a_t) ** 0.5), float(a_t ** 0.5), float(a_p ** 0.5), float(direction), float(noise), float(coeff)]
    return ts + [0], coef


def x0_tables(scheduler, after_skip_steps: int) -> np.ndarray:
    """(sqrt(1 - abar_tt), sqrt(abar_tt)) of every executed step's PREVIOUS timestep tt: the scalars of `reverse_step_pred_x0`
    (text-guided-n-style/inversion/inversion_utils.py:128-140) as called at text-guided-n-style/inversion/h_edit.py:155."""
    ac = scheduler.alphas_cumprod.detach().float().cpu()
    ts = [int(t) for t in scheduler.timesteps[-after_skip_steps:]]
    tts = ts[1:] + [0]
    return np.asarray([[float((1 - ac[tt]) ** 0.5), float(ac[tt] ** 0.5)] for tt in tts], dtype=np.float32)


def skip_pre_coeff(scheduler, after_skip_steps: int, eta, is_ddim_inversion: bool = False):
    """Coefficient of the extra editing move `h_Edit_R_implicit` makes at the first timestep after skipped steps
    (text-guided/inversion/p2p_h_edit.py:214-218,262-263), or None when nothing was skipped."""
    T = scheduler.num_inference_steps
    if after_skip_steps == T:
        return None
    etas = [eta] * T if isinstance(eta, (int, float)) else list(eta)
    ac = scheduler.alphas_cumprod.detach().float().cpu()
    sig, a = (1 - ac) ** 0.5, ac ** 0.5
    ta, t = int(scheduler.timesteps[-(after_skip_steps + 1)]), int(scheduler.timesteps[-after_skip_steps])
    e = etas[after_skip_steps - 1]                       # idx of step i = 0 (p2p_h_edit.py:234)
    omega = 0 if is_ddim_inversion else e * (sig[t] / (sig[ta] * a[t])) * ((ac[t] - ac[ta]) ** 0.5)
    return float((1 - ac[t] - omega ** 2) ** 0.5 - sig[ta] * (a[t] / a[ta]))


class DDIMTables:
    """Minimal DDIM scheduler state for callers that have no diffusers scheduler object (benchmarks, tests):
    scaled-linear betas 0.00085..0.012 over 1000 train steps, set_alpha_to_one=False, "leading" timestep spacing
    with steps_offset (1 for the SD hub config used when eta > 0, 0 for the eta = 0 scheduler built at
    text-guided/main_p2p.py:139-146)."""

    class _Cfg:
        pass

    def __init__(self, num_inference_steps: int, steps_offset: int = 1, num_train_timesteps: int = 1000,
                 beta_start: float = 0.00085, beta_end: float = 0.012):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.config = DDIMTables._Cfg()
        self.config.num_train_timesteps = num_train_timesteps
        self.config.steps_offset = steps_offset
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        """diffusers DDIMScheduler.set_timesteps with timestep_spacing="leading" (what the reference calls at main_p2p.py:148)."""
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        self.timesteps = (torch.arange(0, num_inference_steps) * ratio).round().flip(0).to(torch.int64) + self.config.steps_offset
