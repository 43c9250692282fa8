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
