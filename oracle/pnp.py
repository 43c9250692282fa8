"""ORACLE (test infrastructure, not product code).

CPU restatement of the reference's Plug-and-Play attention/feature injection
(/root/reference/text-guided/plug_n_play/pnp_utils.py) for the oracle SD-1.x UNet, and of the sampler that drives it
(/root/reference/text-guided/inversion/pnp_h_edit.py:33-160).  Pinned against the unmodified reference functions by
tests/golden/*pnp*.pt (tests/make_golden.py) in tests/test_oracle_pin.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Set

import torch

from .h_edit import full_coeff, reverse_step

ATTN_BLOCKS = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}      # pnp_utils.py:88


@dataclass
class PnPState:
    qk_schedule: Set[int] = field(default_factory=set)      # register_attention_control_efficient(model, schedule)  :29
    conv_schedule: Set[int] = field(default_factory=set)    # register_conv_control_efficient(model, schedule)       :97
    t: int = -1                                             # register_time(model, t)                                :12

    def on(self, sched) -> bool:
        return self.t in sched or self.t == 1000            # :51-52, :138


class _PnPSelfAttnProcessor:
    """sa_forward (pnp_utils.py:37-84): for a batch of TWO samples the second takes the first's q and k."""

    def __init__(self, state: PnPState):
        self.state = state

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **_):
        x = hidden_states
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        q, k = attn.to_q(x), attn.to_k(ctx)
        if encoder_hidden_states is None and self.state.on(self.state.qk_schedule):
            src = int(q.shape[0] // 2)
            if src == 1:                                     # :59-63
                q[src:2 * src] = q[:src]
                k[src:2 * src] = k[:src]
        q, k = attn.head_to_batch_dim(q), attn.head_to_batch_dim(k)
        v = attn.head_to_batch_dim(attn.to_v(ctx))
        sim = torch.einsum("b i d, b j d -> b i j", q, k) * attn.scale      # :75
        out = torch.einsum("b i j, b j d -> b i d", sim.softmax(dim=-1), v)
        return attn.to_out[0](attn.batch_to_head_dim(out))


def install_oracle_pnp(unet, state: PnPState) -> None:
    """register_attention_control_efficient + register_conv_control_efficient (pnp_utils.py:29-95,97-164)."""
    for res, blocks in ATTN_BLOCKS.items():
        for b in blocks:
            unet.up_blocks[res].attentions[b].transformer_blocks[0].attn1.set_processor(_PnPSelfAttnProcessor(state))
    block = unet.up_blocks[1].resnets[1]

    def forward(x, temb):                                    # conv_forward (:99-160) for SD-1.x (no up/down-sampling resnets)
        h = block.conv1(torch.nn.functional.silu(block.norm1(x)))
        h = h + block.time_emb_proj(torch.nn.functional.silu(temb))[:, :, None, None]
        h = block.conv2(torch.nn.functional.silu(block.norm2(h)))
        if state.on(state.conv_schedule):
            src = int(h.shape[0] // 2)
            if src == 1:                                     # :144-146
                h[src:2 * src] = h[:src]
        if block.conv_shortcut is not None:
            x = block.conv_shortcut(x)
        return x + h

    block.forward = forward


@torch.no_grad()
def h_edit_pnp_implicit(unet, sched, ctx_uncond, ctx_src, ctx_tar, xT, zs, state: PnPState, cfg_scales: Sequence[float], eta: float = 1.0,
                        optimization_steps: int = 1, after_skip_steps: Optional[int] = None, is_ddim_inversion: bool = False,
                        trace: Optional[List[torch.Tensor]] = None):
    """pnp_h_edit.py:33-160.  `unet` must carry install_oracle_pnp(unet, state).  Returns (edited, reconstructed)."""
    T = sched.num_inference_steps
    S = T if after_skip_steps is None else after_skip_steps
    w_src, w_src_edit, w_tar = [float(c) for c in cfg_scales]
    ab = sched.alphas_cumprod
    op = [int(t) for t in sched.timesteps[-S:]]
    pos = {t: k for k, t in enumerate(op)}
    xt = torch.cat([xT.reshape(1, *xT.shape[-3:])] * 2)
    for i, t in enumerate(op):
        idx = T - pos[t] - (T - S + 1)
        state.t = t                                                                                     # :106
        out = unet(torch.cat([xt, xt]), t, encoder_hidden_states=torch.cat([ctx_uncond, ctx_uncond, ctx_src, ctx_src])).sample   # :117
        e_u, e_c = out.chunk(2)
        prev = reverse_step(sched, e_u + w_src * (e_c - e_u), t, xt, eta, zs[idx], is_ddim_inversion)   # :123
        x_orig, x_base = prev.chunk(2)
        tt = op[i + 1] if i < len(op) - 1 else 0
        x_opt = x_base.clone()
        for _ in range(optimization_steps):
            state.t = tt                                                                                # :140
            c_src = unet(x_opt, tt, encoder_hidden_states=ctx_src).sample                               # :143
            u_tar = unet(x_opt, tt, encoder_hidden_states=ctx_uncond).sample                            # :144
            pair = unet(torch.cat([x_orig, x_opt]), tt, encoder_hidden_states=torch.cat([ctx_src, ctx_tar])).sample   # :150 (PnP)
            c_tar = pair[1:2]
            eps_src_edit = u_tar + w_src_edit * (c_src - u_tar)
            eps_tar = u_tar + w_tar * (c_tar - u_tar)
            coeff = full_coeff(sched, t, tt, eta, is_ddim_inversion) - (1 - ab[t]) ** 0.5 * (ab[tt] ** 0.5 / ab[t] ** 0.5)    # :161-162
            x_opt = x_opt + coeff * (eps_tar - eps_src_edit)                                            # :164-167
        xt = torch.cat([x_orig, x_opt])
        if trace is not None:
            trace.append(xt.clone())
    return xt[1:2].clone(), xt[0:1].clone()
