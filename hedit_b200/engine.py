"""Python handle on the native UNet engine + batched edit loop (C ABI in include/hedit_b200.h).
PyTorch is used only for device memory and streams; all compute happens in libhedit_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .p2p import EditPlan


def unet_config_of(unet) -> dict:
    """Geometry of a diffusers-style (or oracle) SD-1.x UNet object."""
    cfg = getattr(unet, "cfg", None) or getattr(unet, "config", None)
    get = (lambda k, d=None: getattr(cfg, k, d)) if not isinstance(cfg, dict) else (lambda k, d=None: cfg.get(k, d))
    heads = get("attention_head_dim", 8)
    return dict(in_channels=get("in_channels", 4), out_channels=get("out_channels", 4), sample_size=get("sample_size", 64),
                block_out_channels=tuple(get("block_out_channels", (320, 640, 1280, 1280))), layers_per_block=get("layers_per_block", 2),
                heads=heads if isinstance(heads, int) else heads[0], cross_attention_dim=get("cross_attention_dim", 768),
                norm_groups=get("norm_num_groups", 32), ctx_len=77)


class UNetEngine:
    def __init__(self, config: dict, max_samples: int, max_contexts: Optional[int] = None, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("hedit_b200: no CUDA device visible; the B200 path has no CPU fallback")
        c = _lib.UNetConfigC()
        c.in_channels, c.out_channels, c.sample_size = config["in_channels"], config["out_channels"], config["sample_size"]
        for i, v in enumerate(config["block_out_channels"]):
            c.block_out_channels[i] = v
        c.layers_per_block, c.heads = config["layers_per_block"], config["heads"]
        c.cross_attention_dim, c.norm_groups, c.ctx_len = config["cross_attention_dim"], config["norm_groups"], config["ctx_len"]
        self.config = dict(config)
        self.device = device
        self.max_samples = max_samples
        self.max_contexts = max_contexts or max(max_samples, 4)
        self.handle = self.lib.hedit_engine_create(C.byref(c), max_samples, self.max_contexts, device)
        if not self.handle:
            raise RuntimeError("hedit_b200: engine creation failed: " + _lib.last_error())
        self.latent_shape = (config["in_channels"], config["sample_size"], config["sample_size"])
        self.last_stats: Dict[str, int] = {}

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hedit_engine_destroy(h)

    # ---- weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for name, t in sd.items():
            if not torch.is_floating_point(t):
                continue
            t = t.detach().to(torch.float32).contiguous()
            dims = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.hedit_engine_load_tensor(self.handle, name.encode(), t.data_ptr(), dims, t.dim()), f"load {name}")
        _lib.check(self.lib.hedit_engine_finalize(self.handle), "finalize weights")

    @classmethod
    def from_unet(cls, unet, max_samples: int, max_contexts: Optional[int] = None, device: int = 0) -> "UNetEngine":
        eng = cls(unet_config_of(unet), max_samples, max_contexts, device)
        eng.load_state_dict(unet.state_dict())
        return eng

    def tensor_specs(self):
        """[(diffusers parameter name, shape)] the engine expects."""
        out = []
        buf = C.create_string_buffer(256)
        dims = (C.c_int64 * 4)()
        for i in range(self.lib.hedit_engine_tensor_count(self.handle)):
            nd = _lib.check(self.lib.hedit_engine_tensor_info(self.handle, i, buf, 256, dims), "tensor_info")
            out.append((buf.value.decode(), tuple(int(dims[k]) for k in range(nd))))
        return out

    def load_random_weights(self, seed: int = 0) -> None:
        """Synthetic random-init weights generated on the device (benchmarks: no pretrained weights exist offline).
        Variance-preserving scale; norm gammas near 1."""
        g = torch.Generator(device=f"cuda:{self.device}").manual_seed(seed)
        dev = torch.device("cuda", self.device)
        for name, shape in self.tensor_specs():
            if len(shape) >= 2:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * (3.0 / fan_in) ** 0.5
            elif name.endswith("weight"):
                t = 0.8 + 0.4 * torch.rand(shape, generator=g, device=dev)
            else:
                t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * 0.1
            dims = (C.c_int64 * len(shape))(*shape)
            _lib.check(self.lib.hedit_engine_load_tensor(self.handle, name.encode(), t.data_ptr(), dims, len(shape)), f"load {name}")
        _lib.check(self.lib.hedit_engine_finalize(self.handle), "finalize weights")

    def profile_forward(self, S: int, reps: int = 3):
        """{op tag: (ms per forward, launches per forward)} measured with CUDA events around every kernel."""
        buf = C.create_string_buffer(65536)
        _lib.check(self.lib.hedit_engine_profile_forward(self.handle, S, reps, buf, 65536), "profile")
        out, self.last_profile_whole = {}, {}
        for rec in buf.value.decode().split(";"):
            if rec:
                tag, ms, n = rec.split(":")
                if tag.startswith("@"):          # whole forward: launched kernel by kernel / replayed from a CUDA graph
                    self.last_profile_whole[tag[1:]] = float(ms)
                else:
                    out[tag] = (float(ms), int(n))
        return out

    def new_blend_state(self, B: int) -> torch.Tensor:
        """Zeroed LocalBlend accumulator for single-step use (hedit_edit_args.blend_state)."""
        n = self.lib.hedit_engine_blend_state_elems(self.handle, B)
        return torch.zeros(max(n, 1), dtype=torch.float32, device=torch.device("cuda", self.device))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- plain UNet call (use_controller=False)
    def forward(self, x: torch.Tensor, timesteps, ctx: torch.Tensor, ctx_index=None) -> torch.Tensor:
        """eps = unet(x, t, ctx).  ctx is (S,77,D), or (n_ctx,77,D) with ctx_index[S] selecting a context per sample.
        Batches larger than max_samples are processed in chunks."""
        dev = torch.device("cuda", self.device)
        x = x.to(dev, torch.float32).contiguous()
        S = x.shape[0]
        ts = np.ascontiguousarray(np.broadcast_to(np.asarray(timesteps, dtype=np.float32).reshape(-1), (S,)))
        ctx = ctx.to(torch.float32).contiguous()
        eps = torch.empty_like(x)
        if ctx_index is None:
            assert ctx.shape[0] == S
            idx = np.arange(S, dtype=np.int32)
        else:
            idx = np.ascontiguousarray(np.asarray(ctx_index, dtype=np.int32))
        launches = 0
        for lo in range(0, S, self.max_samples):
            hi = min(S, lo + self.max_samples)
            if ctx_index is None:
                c, ci = ctx[lo:hi].contiguous(), np.arange(hi - lo, dtype=np.int32)
            else:
                c, ci = ctx, np.ascontiguousarray(idx[lo:hi])
            assert c.shape[0] <= self.max_contexts, "more contexts than the engine was created for"
            tsc = np.ascontiguousarray(ts[lo:hi])
            launches += _lib.check(self.lib.hedit_unet_forward_indexed(self.handle, x[lo:hi].data_ptr(), tsc.ctypes.data, c.data_ptr(), c.shape[0],
                                                                       ci.ctypes.data, hi - lo, eps[lo:hi].data_ptr(), self._stream()), "unet forward")
        self.last_stats = {"kernel_launches": launches, "sample_forwards": S}
        return eps

    def set_graph_replay(self, on: bool) -> None:
        """CUDA-graph replay of the loop's UNet launches (default on); off = every kernel launched directly.  Bit-identical results."""
        _lib.check(self.lib.hedit_engine_set_graph_replay(self.handle, int(bool(on))), "set_graph_replay")

    def set_prefix_dedup(self, on: bool) -> None:
        """Evaluate the UNet's context-free prefix once per distinct latent of a launch (default on).  Bit-identical results."""
        _lib.check(self.lib.hedit_engine_set_prefix_dedup(self.handle, int(bool(on))), "set_prefix_dedup")

    def set_splitk(self, on: bool) -> None:
        """Split-K for launches with far fewer tiles than SMs (1-5 samples at the deep levels).  Off by default: with it the low bits of a
        result depend on the batch size (results stay run-to-run reproducible); the one-image samplers switch it on."""
        _lib.check(self.lib.hedit_engine_set_splitk(self.handle, int(bool(on))), "set_splitk")

    def n_transformer_blocks(self) -> int:
        """Transformer blocks of the SD-1.x layout (attention on every level but the deepest, plus the mid block): 16 for SD-1.5;
        the reference's controllers count 2 attention layers per block (ptp_utils.py:277-295)."""
        levels, lpb = len(self.config["block_out_channels"]) - 1, self.config["layers_per_block"]
        return levels * lpb + 1 + levels * (lpb + 1)

    # ---- compat path: attention probabilities materialised and handed to a Python hook (include/hedit_b200.h: hedit_unet_forward_compat)
    def forward_compat(self, x: torch.Tensor, timesteps, ctx: torch.Tensor, probs_hook) -> torch.Tensor:
        """eps = unet(x, t, ctx) with `probs_hook(tf_index, is_cross, place, probs)` called for every attention layer in the reference's
        processor order; probs is a (S*heads, n_query, n_key) fp32 CUDA tensor VIEW of the engine's buffer, edited in place
        (p2p/ptp_utils.py:96-107).  place is 0/1/2 = down/mid/up."""
        dev = torch.device("cuda", self.device)
        x = x.to(dev, torch.float32).contiguous()
        S = x.shape[0]
        assert S <= self.max_samples and ctx.shape[0] == S and S <= self.max_contexts
        ts = np.ascontiguousarray(np.broadcast_to(np.asarray(timesteps, dtype=np.float32).reshape(-1), (S,)))
        ctx = ctx.to(dev, torch.float32).contiguous()
        eps = torch.empty_like(x)
        failure = []

        class _View:                      # zero-copy torch view of the device buffer
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 2}

        def _cb(_user, tf_index, is_cross, place, ptr, bh, nq, nk):
            try:
                with torch.cuda.device(dev):
                    probs_hook(tf_index, bool(is_cross), place, torch.as_tensor(_View(ptr, (bh, nq, nk)), device=dev))
                return 0
            except BaseException as ex:   # noqa: BLE001 -- re-raised below, on the caller's side of the C boundary
                failure.append(ex)
                return 1

        cb = _lib.ATTN_PROBS_FN(_cb)
        rc = self.lib.hedit_unet_forward_compat(self.handle, x.data_ptr(), ts.ctypes.data, ctx.data_ptr(), S, eps.data_ptr(), cb, None, self._stream())
        if failure:
            raise failure[0]
        launches = _lib.check(rc, "unet forward (compat)")
        self.last_stats = {"kernel_launches": launches, "sample_forwards": S}
        return eps

    def forward_editor(self, x: torch.Tensor, timesteps, ctx: torch.Tensor, editor_hook) -> torch.Tensor:
        """eps = unet(x, t, ctx) with `editor_hook(tf_index, is_cross, place, q, k, v, sim, attn, heads) -> out` called for every attention
        layer (MasaCtrl's editor protocol, masactrl/masactrl_utils.py:40-89): q/k/v (S*heads, n, d), sim/attn (S*heads, n, m) fp32 CUDA
        views; the returned (S, n, heads*d) tensor becomes the layer's attention output."""
        dev = torch.device("cuda", self.device)
        x = x.to(dev, torch.float32).contiguous()
        S = x.shape[0]
        assert S <= self.max_samples and ctx.shape[0] == S and S <= self.max_contexts
        ts = np.ascontiguousarray(np.broadcast_to(np.asarray(timesteps, dtype=np.float32).reshape(-1), (S,)))
        ctx = ctx.to(dev, torch.float32).contiguous()
        eps = torch.empty_like(x)
        heads = self.config["heads"]
        failure = []

        class _View:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 2}

        def view(ptr, *shape):
            return torch.as_tensor(_View(ptr, shape), device=dev)

        def _cb(_user, tf_index, is_cross, place, q, k, v, sim, attn, out, bh, nq, nk, d):
            try:
                with torch.cuda.device(dev):
                    res = editor_hook(tf_index, bool(is_cross), place, view(q, bh, nq, d), view(k, bh, nk, d), view(v, bh, nk, d),
                                      view(sim, bh, nq, nk), view(attn, bh, nq, nk), heads)
                    view(out, bh // heads, nq, heads * d).copy_(res)
                return 0
            except BaseException as ex:   # noqa: BLE001 -- re-raised below
                failure.append(ex)
                return 1

        cb = _lib.ATTN_EDITOR_FN(_cb)
        rc = self.lib.hedit_unet_forward_editor(self.handle, x.data_ptr(), ts.ctypes.data, ctx.data_ptr(), S, eps.data_ptr(), cb, None, self._stream())
        if failure:
            raise failure[0]
        launches = _lib.check(rc, "unet forward (editor)")
        self.last_stats = {"kernel_launches": launches, "sample_forwards": S}
        return eps

    # ---- the bridge-sampling loop
    def edit(self, xT: torch.Tensor, zs: torch.Tensor, ctx: torch.Tensor, timesteps: Sequence[int], coef: np.ndarray,
             cfg_scales: Sequence[float], plan: Optional[EditPlan], weight_reconstruction: float = 0.1, optimization_steps: int = 1,
             explicit_form: bool = False, schedule: int = 1, trace: bool = False, variant: int = 0, masactrl=None, mos_pull: bool = True,
             xt_is_pair: bool = False, ctrl_step0: int = 0, blend_state: Optional[torch.Tensor] = None, pnp=None, pre_coeff=None, guidance=None,
             coef_edit: Optional[np.ndarray] = None):
        """xT (B,C,h,w), zs (B,steps,C,h,w), ctx (1+2B,77,D): all on the SAME side (all host or all on this device).
        variant 1 = h_Edit_R_* (no attention control); variant 2 = the EF / PnP-Inversion baseline samplers (coef_edit = reverse-step
        scalars of the edit row when they differ from the orig row's); masactrl = (layer_mask, step_on[steps*K]) enables mutual self-attention in the
        transformer blocks of layer_mask during the controlled launches flagged in step_on;
        pnp = (self_mask, qk_on[steps], feat_on[steps]) runs h_Edit_PnP_implicit (Plug-and-Play q/k and feature injection);
        guidance = (fn, weight, x0_coef[steps,2]) adds the reward-guided Langevin move: fn(x0 (B,C,h,w) cuda tensor) -> dLoss/dx0.
        Returns (edited, recon[, trace]) on that side."""
        B, steps = xT.shape[0], zs.shape[1]
        out_shape = (B,) + tuple(xT.shape[-3:])
        on_host = not xT.is_cuda
        assert zs.is_cuda == xT.is_cuda, "xT and zs must live on the same side"
        xT = xT.to(torch.float32).contiguous()
        zs = zs.to(torch.float32).contiguous()
        ctx = ctx.to(torch.float32).contiguous()
        assert ctx.shape[0] == 1 + 2 * B and zs.shape[0] == B and len(timesteps) == steps + 1 and coef.shape == (steps, 6)
        mk = (lambda *s: torch.empty(*s, dtype=torch.float32, pin_memory=True)) if on_host else \
            (lambda *s: torch.empty(*s, dtype=torch.float32, device=xT.device))
        edited, recon = mk(*out_shape), mk(*out_shape)
        tr = mk(steps, B, 2, *out_shape[1:]) if trace else None
        ts = np.ascontiguousarray(np.asarray(timesteps, dtype=np.float32))
        coef = np.ascontiguousarray(coef, dtype=np.float32)
        a = _lib.EditArgsC()
        a.B, a.steps, a.opt_steps, a.explicit_form, a.schedule, a.buffers_on_host = B, steps, optimization_steps, int(explicit_form), schedule, int(on_host)
        a.xT, a.zs, a.ctx, a.timesteps, a.coef = xT.data_ptr(), zs.data_ptr(), ctx.data_ptr(), ts.ctypes.data, coef.ctypes.data
        a.w_src, a.w_src_edit, a.w_tar = [float(v) for v in cfg_scales]
        a.weight_reconstruction = float(weight_reconstruction)
        a.variant = int(variant)
        a.mos_pull = int(mos_pull)
        keep = []
        if coef_edit is not None:
            coef_edit = np.ascontiguousarray(coef_edit, dtype=np.float32)
            assert coef_edit.shape == (steps, 6)
            keep.append(coef_edit)
            a.coef_edit = coef_edit.ctypes.data
        if masactrl is not None:
            n_ctrl = steps * (1 if explicit_form else max(1, optimization_steps))
            step_on = np.ascontiguousarray(np.asarray(masactrl[1], dtype=np.int32))
            assert step_on.shape == (n_ctrl,), f"masactrl step flags must cover the {n_ctrl} controlled launches"
            keep.append(step_on)
            a.masa, a.masa_layer_mask, a.masa_step_on = 1, int(masactrl[0]), step_on.ctypes.data
        a.xt_is_pair, a.ctrl_step0 = int(xt_is_pair), int(ctrl_step0)
        a.blend_state = blend_state.data_ptr() if blend_state is not None else None
        keep += [ts, coef, xT, zs, ctx]
        cb_error = []
        if guidance is not None:
            fn, weight, x0_coef = guidance
            dev = torch.device("cuda", self.device)
            x0_buf = torch.empty(out_shape, dtype=torch.float32, device=dev)
            grad_buf = torch.zeros(out_shape, dtype=torch.float32, device=dev)
            x0_coef = np.ascontiguousarray(x0_coef, dtype=np.float32)
            assert x0_coef.shape == (steps, 2)

            def _cb(_user, step, opt_step):
                try:       # runs on the launching stream (torch's current stream): x0_buf is ready in stream order
                    g = fn(x0_buf)
                    grad_buf.copy_(g.reshape(out_shape).to(torch.float32))
                    return 0
                except Exception as exc:       # surfaced after the native call returns
                    cb_error.append(exc)
                    return -1

            cfn = _lib.GUIDANCE_FN(_cb)
            keep += [cfn, x0_buf, grad_buf, x0_coef]
            a.guidance = C.cast(cfn, C.c_void_p)
            a.guidance_weight, a.x0_coef = float(weight), x0_coef.ctypes.data
            a.guid_x0, a.guid_grad = x0_buf.data_ptr(), grad_buf.data_ptr()
        if pre_coeff is not None:      # h_Edit_R_implicit on a skipped schedule
            a.pre_step, a.pre_coeff = 1, float(pre_coeff)
        if pnp is not None:
            qk_on, feat_on = (np.ascontiguousarray(np.asarray(v, dtype=np.int32)) for v in pnp[1:])
            assert qk_on.shape == (steps,) and feat_on.shape == (steps,)
            keep += [qk_on, feat_on]
            a.pnp, a.pnp_self_mask, a.pnp_qk_on, a.pnp_feat_on = 1, int(pnp[0]), qk_on.ctypes.data, feat_on.ctypes.data
        if plan is not None:
            a.use_p2p = 1
            arrs = dict(mapper=plan.mapper, is_replace=plan.is_replace, c_base=plan.c_base, c_tar=plan.c_tar,
                        has_blend=plan.has_blend, blend_alpha=plan.blend_alpha)
            for k, v in arrs.items():
                v = np.ascontiguousarray(v)
                keep.append(v)
                setattr(a, k, v.ctypes.data)
            if plan.replace_m is not None:
                rm = np.ascontiguousarray(plan.replace_m)
                keep.append(rm)
                a.replace_m = rm.ctypes.data
            if getattr(plan, "map_w", None) is not None:
                mw = np.ascontiguousarray(plan.map_w, dtype=np.float32)
                keep.append(mw)
                a.map_w, a.map_rows = mw.ctypes.data, int(mw.shape[1])
            if not plan.has_blend.any():
                a.has_blend = None
            a.self_lo, a.self_hi = plan.self_window
            a.self_max_tokens = 32 * 32
            a.start_blend, a.blend_th = plan.start_blend, plan.blend_th
            a.blend_rows, a.blend_th_sub = int(plan.blend_alpha.shape[1]), float(getattr(plan, "blend_th_sub", plan.blend_th))
        a.edited, a.recon = edited.data_ptr(), recon.data_ptr()
        a.trace = tr.data_ptr() if tr is not None else None
        rc = self.lib.hedit_edit_p2p(self.handle, C.byref(a), self._stream())
        if cb_error:
            raise cb_error[0]
        _lib.check(rc, "edit")
        self.last_stats = {"sample_forwards": int(a.n_sample_forwards), "kernel_launches": int(a.n_kernel_launches)}
        return (edited, recon, tr) if trace else (edited, recon)
