"""Host-side Prompt-to-Prompt set-up for the B200 path: builds the per-image controller tables and compiles them
into the per-step device "edit plan" the fused cross-/self-attention kernels consume.

Mirrors the reference's set-up surface so a driver can switch imports:
  make_controller(...)            <- text-guided/p2p/ptp_controller_utils.py:106
  register_attention_control(...) <- text-guided/p2p/ptp_utils.py:277
  EditController attributes       <- AttentionControlEdit / Refine / Replace / Reweight / LocalBlend
                                     (text-guided/p2p/ptp_classes.py:17-283)
`compile_edit_plan` also accepts the reference's own controller objects (duck-typed attributes), so a controller
built by the reference code drops straight into the fast path.

Generalisation over the reference: tables carry a leading image dimension B (the reference hard-codes one image,
ptp_classes.py:178,209-213).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

MAX_WORDS = 77
PAD = 80          # token axis padded to a multiple of 16 for the tensor-core tiles


# ----------------------------------------------------------------------------------------------------------------
# token bookkeeping
# ----------------------------------------------------------------------------------------------------------------
def get_word_inds(text: str, word_place: Union[int, str], tokenizer) -> np.ndarray:
    """Token positions (BOS-shifted) covered by a word of `text` (ptp_utils.py:297-315 semantics)."""
    words = text.split(" ")
    if isinstance(word_place, str):
        wanted = {i for i, w in enumerate(words) if w == word_place}
    else:
        wanted = {int(word_place)}
    if not wanted:
        return np.zeros(0, dtype=np.int64)
    ids = tokenizer.encode(text)[1:-1]
    lens = [len(tokenizer.decode([t]).strip("#")) for t in ids]
    out, w, filled = [], 0, 0
    for pos, ln in enumerate(lens):
        filled += ln
        if w in wanted:
            out.append(pos + 1)
        if filled >= len(words[w]):
            w, filled = w + 1, 0
    return np.asarray(out, dtype=np.int64)


def _align_tokens(x: Sequence[int], y: Sequence[int]) -> np.ndarray:
    """Global alignment (match +1, mismatch -1, gap 0; ties prefer gap-in-x, then gap-in-y, then diagonal --
    seq_aligner.py:60-109).  Returns for every position of y the aligned position of x, or -1."""
    nx, ny = len(x), len(y)
    sc = np.zeros((nx + 1, ny + 1), dtype=np.int64)
    tb = np.zeros((nx + 1, ny + 1), dtype=np.int8)
    tb[0, 1:], tb[1:, 0], tb[0, 0] = 1, 2, 4
    xa, ya = np.asarray(x), np.asarray(y)
    for i in range(1, nx + 1):
        match = np.where(xa[i - 1] == ya, 1, -1)
        for j in range(1, ny + 1):
            cand = (sc[i, j - 1], sc[i - 1, j], sc[i - 1, j - 1] + match[j - 1])
            best = max(cand)
            sc[i, j] = best
            tb[i, j] = 1 if cand[0] == best else (2 if cand[1] == best else 3)
    out = np.full(ny, -1, dtype=np.int64)
    i, j = nx, ny
    while i > 0 or j > 0:
        m = tb[i, j]
        if m == 3:
            i, j = i - 1, j - 1
            out[j] = i
        elif m == 1:
            j -= 1
        elif m == 2:
            i -= 1
        else:
            break
    return out


def get_refinement_mapper(prompts: Sequence[str], tokenizer, max_len: int = MAX_WORDS):
    """(mapper[1,77] int64, alphas[1,77]) for prompts = [src, tar] (seq_aligner.py:112-133)."""
    xs, ys = tokenizer.encode(prompts[0]), tokenizer.encode(prompts[1])
    al = _align_tokens(xs, ys)
    mapper = np.zeros(max_len, dtype=np.int64)
    alphas = np.ones(max_len, dtype=np.float32)
    n = min(len(al), max_len)
    mapper[:n] = al[:n]
    alphas[:n] = (al[:n] != -1).astype(np.float32)
    mapper[len(ys):] = len(ys) + np.arange(max_len - len(ys))
    return torch.from_numpy(mapper)[None], torch.from_numpy(alphas)[None]


def get_replacement_mapper(prompts: Sequence[str], tokenizer, max_len: int = MAX_WORDS) -> torch.Tensor:
    """(1,77,77) word-swap matrix (seq_aligner.py:157-199)."""
    wx, wy = prompts[0].split(" "), prompts[1].split(" ")
    if len(wx) != len(wy):
        raise ValueError("attention replacement edit can only be applied on prompts with the same number of words")
    changed = [k for k in range(len(wy)) if wx[k] != wy[k]]
    src = [get_word_inds(prompts[0], k, tokenizer) for k in changed]
    tar = [get_word_inds(prompts[1], k, tokenizer) for k in changed]
    m = np.zeros((max_len, max_len), dtype=np.float32)
    i = j = nxt = 0
    while i < max_len and j < max_len:
        if nxt < len(src) and src[nxt][0] == i:
            s, t = src[nxt], tar[nxt]
            if len(s) == len(t):
                m[s, t] = 1.0
            else:
                m[np.ix_(s, t)] = 1.0 / len(t)
            nxt += 1
            i, j = i + len(s), j + len(t)
        else:
            m[(i if nxt < len(src) else j), j] = 1.0
            i, j = i + 1, j + 1
    return torch.from_numpy(m)[None]


def get_time_words_attention_alpha(prompts, num_steps, cross_replace_steps, tokenizer, max_num_words=MAX_WORDS):
    """(T+1, 1, 1, 1, 77) 0/1 schedule of the cross-attention injection (ptp_utils.py:318-349)."""
    spec = dict(cross_replace_steps) if isinstance(cross_replace_steps, dict) else {"default_": cross_replace_steps}
    spec.setdefault("default_", (0.0, 1.0))
    rows = num_steps + 1
    tab = np.zeros((rows, len(prompts) - 1, max_num_words), dtype=np.float32)

    def window(b):
        lo, hi = (0.0, b) if isinstance(b, float) else b
        return int(lo * rows), int(hi * rows)

    a, b = window(spec["default_"])
    tab[a:b] = 1.0
    for word, bounds in spec.items():
        if word == "default_":
            continue
        a, b = window(bounds)
        for p in range(1, len(prompts)):
            cols = get_word_inds(prompts[p], word, tokenizer)
            if len(cols):
                tab[:, p - 1, cols] = 0.0
                tab[a:b, p - 1, cols] = 1.0
    return torch.from_numpy(tab).reshape(rows, len(prompts) - 1, 1, 1, max_num_words)


def get_equalizer(text: str, word_select, values, tokenizer) -> torch.Tensor:
    """ones(1,77) with per-word multipliers (ptp_controller_utils.py:92-104)."""
    if isinstance(word_select, (int, str)):
        word_select = (word_select,)
    eq = torch.ones(1, MAX_WORDS)
    for w, v in zip(word_select, values):
        eq[:, torch.as_tensor(get_word_inds(text, w, tokenizer), dtype=torch.int64)] = float(v)
    return eq


# ----------------------------------------------------------------------------------------------------------------
# controller objects (attribute-compatible with the reference classes)
# ----------------------------------------------------------------------------------------------------------------
class LocalBlend:
    """Holds LocalBlend's constants (ptp_classes.py:17-42); the mask itself is computed on the GPU by
    local_blend_kernel from the accumulated 16x16 cross-attention word maps."""

    def __init__(self, prompts, num_steps, words, start_blend=0.2, th=(0.3, 0.3), tokenizer=None, device=None):
        al = torch.zeros(len(prompts), 1, 1, 1, 1, MAX_WORDS)
        for r, (prompt, ws) in enumerate(zip(prompts, words)):
            for w in ([ws] if isinstance(ws, str) else ws):
                al[r, :, :, :, :, torch.as_tensor(get_word_inds(prompt, w, tokenizer), dtype=torch.int64)] = 1
        self.alpha_layers = al
        self.substruct_layers = None
        self.start_blend = int(start_blend * num_steps)
        self.counter = 0
        self.th = th


class EditController:
    """One image's P2P controller: Refine or Replace, optional Reweight equalizer, optional LocalBlend.
    Carries the reference controllers' observable attributes (cur_step, cur_att_layer, num_att_layers, ...)."""

    def __init__(self, prompts, num_steps, cross_replace_steps, self_replace_steps, is_replace, local_blend=None,
                 equalizer=None, tokenizer=None, device=None):
        self.prompts = list(prompts)
        self.num_steps = num_steps
        self.batch_size = len(prompts)
        self.cross_replace_alpha = get_time_words_attention_alpha(prompts, num_steps, cross_replace_steps, tokenizer)
        sa = (0.0, self_replace_steps) if isinstance(self_replace_steps, float) else self_replace_steps
        self.num_self_replace = (int(num_steps * sa[0]), int(num_steps * sa[1]))
        self.local_blend = local_blend
        self.is_replace = bool(is_replace)
        if self.is_replace:
            self.mapper = get_replacement_mapper(prompts, tokenizer)          # (1,77,77)
            self.alphas = None
        else:
            self.mapper, al = get_refinement_mapper(prompts, tokenizer)       # (1,77)
            self.alphas = al.reshape(al.shape[0], 1, 1, al.shape[1])
        self.equalizer = equalizer                                            # (1,77) or None
        self.prev_controller = None
        self.num_att_layers = -1
        self.cur_step = 0
        self.cur_att_layer = 0
        self.attention_store: Dict[str, list] = {}

    # reference-compatible no-ops on the fast path (state is advanced by the fused kernels / the sampler)
    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0
        self.attention_store = {}


def make_controller(prompts: List[str], is_replace_controller: bool, cross_replace_steps, self_replace_steps,
                    blend_word=None, equilizer_params=None, num_steps=None, tokenizer=None, device=None) -> EditController:
    """Same signature as the reference's make_controller (ptp_controller_utils.py:106-133)."""
    lb = LocalBlend(prompts, num_steps, blend_word, tokenizer=tokenizer, device=device) if blend_word is not None else None
    eq = None
    if equilizer_params is not None:
        eq = get_equalizer(prompts[1], equilizer_params["words"], equilizer_params["values"], tokenizer)
    return EditController(prompts, num_steps, cross_replace_steps, self_replace_steps, is_replace_controller, lb, eq,
                          tokenizer, device)


def register_attention_control(model, controller) -> None:
    """The fused kernels need no per-layer Python processors; this records the layer count the reference stores
    on the controller (ptp_utils.py:277-295)."""
    n = len(getattr(model.unet, "attn_processors", {})) if hasattr(model, "unet") else 0
    controller.num_att_layers = n if n else 32


# ----------------------------------------------------------------------------------------------------------------
# edit-plan compilation
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class EditPlan:
    """Per-batch device tables (host numpy, uploaded by the C ABI)."""
    B: int
    steps: int
    mapper: np.ndarray        # (B,80) int32, clamped
    is_replace: np.ndarray    # (B,) int32
    replace_m: Optional[np.ndarray]   # (B,77,80) float32 or None
    c_base: np.ndarray        # (steps+1,B,80)
    c_tar: np.ndarray         # (steps+1,B,80)
    self_window: Tuple[int, int]
    has_blend: np.ndarray     # (B,) int32
    blend_alpha: np.ndarray   # (B,2,80)
    start_blend: int
    blend_th: float


def _flatten(ctrl):
    """(mapper, refine_alphas, equalizer) from our EditController or the reference's Reweight->Refine/Replace chain."""
    eq = getattr(ctrl, "equalizer", None)
    inner = getattr(ctrl, "prev_controller", None)
    core = inner if inner is not None else ctrl
    mapper = getattr(core, "mapper", None)
    alphas = getattr(core, "alphas", None)
    return mapper, alphas, eq


def compile_edit_plan(controllers: Sequence, steps: int) -> EditPlan:
    B = len(controllers)
    mapper = np.zeros((B, PAD), dtype=np.int32)
    is_rep = np.zeros(B, dtype=np.int32)
    rep_m = None
    c_base = np.zeros((steps + 1, B, PAD), dtype=np.float32)
    c_tar = np.zeros((steps + 1, B, PAD), dtype=np.float32)
    has_blend = np.zeros(B, dtype=np.int32)
    blend_alpha = np.zeros((B, 2, PAD), dtype=np.float32)
    windows, starts, ths = set(), set(), set()
    for b, c in enumerate(controllers):
        aw = torch.as_tensor(c.cross_replace_alpha).reshape(-1, MAX_WORDS).float().cpu().numpy()   # (T+1,77)
        if aw.shape[0] != steps + 1:
            raise ValueError(f"controller {b} was built for {aw.shape[0] - 1} steps, sampler runs {steps}")
        m, ra, eq = _flatten(c)
        eqv = np.ones(MAX_WORDS, dtype=np.float32) if eq is None else torch.as_tensor(eq).reshape(MAX_WORDS).float().cpu().numpy()
        if m is None:                         # bare Reweight: base passes through unchanged
            mv, rav = np.arange(MAX_WORDS), np.ones(MAX_WORDS, dtype=np.float32)
        else:
            m = torch.as_tensor(m).cpu()
            if m.dim() == 3:                  # AttentionReplace (1,77,77)
                is_rep[b] = 1
                if rep_m is None:
                    rep_m = np.zeros((B, MAX_WORDS, PAD), dtype=np.float32)
                rep_m[b, :, :MAX_WORDS] = m[0].float().numpy()
                mv, rav = np.arange(MAX_WORDS), np.ones(MAX_WORDS, dtype=np.float32)
            else:
                mv = m.reshape(MAX_WORDS).numpy()
                rav = torch.as_tensor(ra).reshape(MAX_WORDS).float().cpu().numpy()
        mapper[b, :MAX_WORDS] = np.mod(mv, MAX_WORDS)            # -1 -> 76 like torch indexing; its alpha is 0
        c_base[:, b, :MAX_WORDS] = rav[None] * eqv[None] * aw
        c_tar[:, b, :MAX_WORDS] = (1.0 - rav)[None] * eqv[None] * aw + (1.0 - aw)
        windows.add(tuple(int(v) for v in c.num_self_replace))
        lb = getattr(c, "local_blend", None)
        if lb is not None:
            if getattr(lb, "substruct_layers", None) is not None:
                raise NotImplementedError("LocalBlend substruct_words are not supported on the fused path")
            has_blend[b] = 1
            blend_alpha[b, :, :MAX_WORDS] = torch.as_tensor(lb.alpha_layers).reshape(2, MAX_WORDS).float().cpu().numpy()
            starts.add(int(lb.start_blend))
            ths.add(float(lb.th[0]))
    if len(windows) > 1 or len(starts) > 1 or len(ths) > 1:
        raise ValueError("all images of one batch must share the self-replace window and LocalBlend start/threshold")
    return EditPlan(B=B, steps=steps, mapper=mapper, is_replace=is_rep, replace_m=rep_m, c_base=c_base, c_tar=c_tar,
                    self_window=next(iter(windows)) if windows else (0, 0), has_blend=has_blend, blend_alpha=blend_alpha,
                    start_blend=next(iter(starts)) if starts else 0, blend_th=next(iter(ths)) if ths else 0.3)
