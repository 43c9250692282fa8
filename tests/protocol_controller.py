"""A USER-SIDE controller that relies on nothing but the reference's call protocol: `controller(probs, is_cross, place,
save_attn)` edits the (batch*heads, N, M) probabilities in place and counts layers; `step_callback(x_t)` blends latents.  It is NOT one
of the stock classes the fast path recognises (its class name is its own), so the samplers must serve it through the compat path.
Semantics written from the protocol description of text-guided/p2p/ptp_classes.py:91-107 (call / counters), :135-152 (store),
:202-232 (cross / self replacement), :241-282 (replace / refine / reweight) and :44-72 (LocalBlend); used by tests/test_gpu_compat.py
to check the compat path against the goldens the unmodified reference produced."""
import torch
import torch.nn.functional as F


class UserController:
    def __init__(self, tables, device):
        """tables: an hedit_b200 EditController, used only as a bag of precomputed tensors (alpha schedule, mapper, equalizer, blend words)."""
        self.device = device
        self.alpha = tables.cross_replace_alpha.to(device)                 # (T+1, 1, 1, 1, 77)
        self.self_window = tables.num_self_replace
        self.is_replace = tables.is_replace
        self.mapper = tables.mapper.to(device)
        self.refine_alphas = None if tables.alphas is None else tables.alphas.to(device)
        self.equalizer = None if tables.equalizer is None else tables.equalizer.to(device)
        self.blend = tables.local_blend
        self.blend_calls = 0
        self.n_prompts = 2
        self.num_att_layers = -1
        self.cur_step = 0
        self.cur_att_layer = 0
        self.step_store = self._empty()
        self.attention_store = {}
        self.calls = 0

    @staticmethod
    def _empty():
        return {f"{p}_{k}": [] for p in ("down", "mid", "up") for k in ("cross", "self")}

    # ---- the protocol
    def __call__(self, attn, is_cross, place_in_unet, save_attn):
        self.calls += 1
        half = attn.shape[0] // 2                          # the unconditional half is left alone
        self._edit(attn[half:], is_cross, place_in_unet, save_attn)
        if not save_attn:
            return attn
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self._fold_step()
        return attn

    def step_callback(self, x_t):
        if self.blend is None:
            return x_t
        self.blend_calls += 1
        if self.blend_calls <= self.blend.start_blend:
            return x_t
        maps = self.attention_store["down_cross"][2:4] + self.attention_store["up_cross"][:3]
        al = self.blend.alpha_layers.to(x_t.device)
        maps = torch.cat([m.reshape(al.shape[0], -1, 1, 16, 16, al.shape[-1]) for m in maps], dim=1)
        m = (maps * al).sum(-1).mean(1)
        m = F.max_pool2d(m, (3, 3), (1, 1), padding=(1, 1))
        m = F.interpolate(m, size=x_t.shape[2:])
        m = m / m.amax(dim=(2, 3), keepdim=True)
        m = m.gt(self.blend.th[0])
        m = (m[:1] + m).float()
        return x_t[:1] + m * (x_t - x_t[:1])

    # ---- internals
    def _fold_step(self):
        if not self.attention_store:
            self.attention_store = self.step_store
        else:
            for key, items in self.attention_store.items():
                for i in range(len(items)):
                    items[i] += self.step_store[key][i]
        self.step_store = self._empty()

    def _edit(self, attn, is_cross, place, save_attn):
        if attn.shape[1] <= 32 ** 2 and save_attn:
            # the engine reuses its probabilities buffer for the next layer, so keep a copy (the reference keeps a view of a tensor
            # that is never overwritten); taken AFTER the edit below, which is what the reference's aliasing view ends up holding
            store_after = True
        else:
            store_after = False
        if is_cross or self.self_window[0] <= self.cur_step < self.self_window[1]:
            heads = attn.shape[0] // self.n_prompts
            v = attn.view(self.n_prompts, heads, *attn.shape[1:])
            base, tar = v[0], v[1:]
            if is_cross:
                a = self.alpha[self.cur_step]
                if self.is_replace:
                    mapped = torch.einsum("hpw,bwn->bhpn", base, self.mapper)
                else:
                    mapped = base[:, :, self.mapper].permute(2, 0, 1, 3) * self.refine_alphas + tar * (1 - self.refine_alphas)
                if self.equalizer is not None:
                    mapped = mapped * self.equalizer[:, None, None, :]
                v[1:] = mapped * a + (1 - a) * tar
            elif tar.shape[2] <= 32 ** 2:
                v[1:] = base.unsqueeze(0).expand(tar.shape[0], *base.shape)
        if store_after:
            self.step_store[f"{place}_{'cross' if is_cross else 'self'}"].append(attn.clone())


class UserMutualSelfAttention:
    """A USER-SIDE MasaCtrl editor that relies on nothing but the editor call protocol (text-guided/masactrl/masactrl_utils.py:15-23, 72-75):
    `out = editor(q, k, v, sim, attn, is_cross, place_in_unet, num_heads, scale=...)` with q/k/v (b*h, n, d), returning (b, n, h*d), and
    counting layers itself.  From `start_step` on and in transformer blocks >= `start_layer`, both samples of the unconditional pair and
    of the conditional pair attend to the keys / values of the FIRST sample of their pair (masactrl/masactrl.py:38-64).  Its class name
    is its own, so the sampler must serve it through the compat path."""

    def __init__(self, start_step, start_layer):
        self.start_step, self.start_layer = start_step, start_layer
        self.cur_step = 0
        self.cur_att_layer = 0
        self.num_att_layers = -1
        self.calls = 0
        self.controlled = 0

    def __call__(self, q, k, v, sim, attn, is_cross, place_in_unet, num_heads, **kwargs):
        self.calls += 1
        out = self._forward(q, k, v, attn, is_cross, num_heads, kwargs.get("scale"))
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
        return out

    @staticmethod
    def _merge(x, heads):                        # (b*h, n, d) -> (b, n, h*d)
        bh, n, d = x.shape
        return x.view(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, heads * d)

    def _forward(self, q, k, v, attn, is_cross, heads, scale):
        if is_cross or self.cur_step < self.start_step or self.cur_att_layer // 2 < self.start_layer:
            return self._merge(torch.bmm(attn, v), heads)
        self.controlled += 1
        outs = []
        for qp, kp, vp in zip(q.chunk(2), k.chunk(2), v.chunk(2)):          # unconditional pair, conditional pair
            n, d = qp.shape[1], qp.shape[2]
            qq = qp.view(-1, heads, n, d)
            k1, v1 = kp[:heads], vp[:heads]                                  # the pair's first (source) sample
            p = torch.softmax(torch.einsum("bhid,hjd->bhij", qq, k1) * scale, dim=-1)
            o = torch.einsum("bhij,hjd->bhid", p, v1)
            outs.append(o.permute(0, 2, 1, 3).reshape(qq.shape[0], n, heads * d))
        return torch.cat(outs, dim=0)
