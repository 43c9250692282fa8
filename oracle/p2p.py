"""ORACLE (test infrastructure, not product code).

CPU restatement of the reference's Prompt-to-Prompt attention control for ONE image (prompts = [src, tar]):
set-up tables, the per-layer hook, the per-step map accumulation and LocalBlend.  Each function cites the
reference lines it follows (paths relative to /root/reference/text-guided/).  Pinned against the reference's own
classes by tests/test_oracle_pin.py (run where /root/reference exists) and by tests/golden/*.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

MAX_WORDS = 77


# ----------------------------------------------------------------------------------------------------------
# set-up (host side, once per image)
# ----------------------------------------------------------------------------------------------------------
def word_token_indices(text: str, word_place: Union[int, str], tokenizer) -> np.ndarray:
    """p2p/ptp_utils.py:297-315 -- greedy char-length walk over the decoded tokens; +1 for BOS."""
    words = text.split(" ")
    if isinstance(word_place, str):
        places = [i for i, w in enumerate(words) if w == word_place]
    else:
        places = [int(word_place)]
    hits: List[int] = []
    if places:
        pieces = [tokenizer.decode([tok]).strip("#") for tok in tokenizer.encode(text)][1:-1]
        acc, ptr = 0, 0
        for n, piece in enumerate(pieces):
            acc += len(piece)
            if ptr in places:
                hits.append(n + 1)
            if acc >= len(words[ptr]):
                ptr, acc = ptr + 1, 0
    return np.array(hits, dtype=np.int64)


def cross_alpha_table(prompts: Sequence[str], num_steps: int, xa, tokenizer) -> torch.Tensor:
    """p2p/ptp_utils.py:318-349 -- (T+1, 77) table for the single target prompt; 1 inside [lo, hi)*(T+1)."""
    spec = dict(xa) if isinstance(xa, dict) else {"default_": xa}
    spec.setdefault("default_", (0.0, 1.0))
    rows = num_steps + 1
    table = torch.zeros(rows, MAX_WORDS)

    def stamp(bounds, cols):
        lo, hi = (0, bounds) if isinstance(bounds, float) else bounds
        a, b = int(lo * rows), int(hi * rows)
        table[:a, cols] = 0
        table[a:b, cols] = 1
        table[b:, cols] = 0

    stamp(spec["default_"], slice(None))
    for word, bounds in spec.items():
        if word == "default_":
            continue
        cols = word_token_indices(prompts[1], word, tokenizer)
        if len(cols):
            stamp(bounds, torch.as_tensor(cols))
    return table


def _needleman_wunsch(x: Sequence[int], y: Sequence[int]) -> List[Tuple[int, int]]:
    """p2p/seq_aligner.py:60-109 -- gap 0, match +1, mismatch -1; tie-break left > up > diag.
    Returns for each y position its aligned x position or -1."""
    nx, ny = len(x), len(y)
    score = np.zeros((nx + 1, ny + 1), dtype=np.int32)      # gap = 0 -> zero borders
    move = np.zeros((nx + 1, ny + 1), dtype=np.int8)
    move[0, 1:], move[1:, 0], move[0, 0] = 1, 2, 4
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            left, up = score[i, j - 1], score[i - 1, j]
            diag = score[i - 1, j - 1] + (1 if x[i - 1] == y[j - 1] else -1)
            best = max(left, up, diag)
            score[i, j] = best
            move[i, j] = 1 if best == left else (2 if best == up else 3)
    i, j, pairs = nx, ny, []
    while i > 0 or j > 0:
        m = move[i, j]
        if m == 3:
            i, j = i - 1, j - 1
            pairs.append((j, i))
        elif m == 1:
            j -= 1
            pairs.append((j, -1))
        elif m == 2:
            i -= 1
        else:
            break
    pairs.reverse()
    return pairs


def refinement_mapper(prompts: Sequence[str], tokenizer) -> Tuple[torch.Tensor, torch.Tensor]:
    """p2p/seq_aligner.py:112-133 -- (mapper[77] int64, alphas[77])."""
    xs, ys = tokenizer.encode(prompts[0]), tokenizer.encode(prompts[1])
    pairs = _needleman_wunsch(xs, ys)
    src = torch.tensor([p[1] for p in pairs], dtype=torch.int64)
    n = src.shape[0]
    alphas = torch.ones(MAX_WORDS)
    alphas[:n] = (src != -1).float()
    mapper = torch.zeros(MAX_WORDS, dtype=torch.int64)
    mapper[:n] = src
    mapper[n:] = len(ys) + torch.arange(MAX_WORDS - len(ys))
    return mapper, alphas


def replacement_mapper(prompts: Sequence[str], tokenizer) -> torch.Tensor:
    """p2p/seq_aligner.py:157-199 -- (77,77) float matrix, identity off the replaced words."""
    wx, wy = prompts[0].split(" "), prompts[1].split(" ")
    if len(wx) != len(wy):
        raise ValueError("attention replacement needs prompts with the same number of words")
    diff = [i for i in range(len(wy)) if wy[i] != wx[i]]
    src_ix = [word_token_indices(prompts[0], i, tokenizer) for i in diff]
    tar_ix = [word_token_indices(prompts[1], i, tokenizer) for i in diff]
    m = np.zeros((MAX_WORDS, MAX_WORDS))
    i = j = cur = 0
    while i < MAX_WORDS and j < MAX_WORDS:
        if cur < len(src_ix) and src_ix[cur][0] == i:
            s, t = src_ix[cur], tar_ix[cur]
            if len(s) == len(t):
                m[s, t] = 1
            else:
                for tt in t:
                    m[s, tt] = 1.0 / len(t)
            cur += 1
            i += len(s)
            j += len(t)
        elif cur < len(src_ix):
            m[i, j] = 1
            i, j = i + 1, j + 1
        else:
            m[j, j] = 1
            i, j = i + 1, j + 1
    return torch.from_numpy(m).float()


def equalizer_row(text: str, words, values, tokenizer) -> torch.Tensor:
    """p2p/ptp_controller_utils.py:92-104 -- ones(77) with `value` at the word's tokens."""
    if isinstance(words, (int, str)):
        words = (words,)
    eq = torch.ones(MAX_WORDS)
    for w, v in zip(words, values):
        eq[torch.as_tensor(word_token_indices(text, w, tokenizer))] = v
    return eq


@dataclass
class EditSpec:
    """Everything `make_controller` (p2p/ptp_controller_utils.py:106-133) bakes into the controller objects."""
    num_steps: int
    is_replace: bool
    alpha_words: torch.Tensor                      # (T+1, 77)      AttentionControlEdit.cross_replace_alpha
    self_window: Tuple[int, int]                   #                 .num_self_replace
    mapper: Optional[torch.Tensor] = None          # (77,) int64     AttentionRefine.mapper
    refine_alpha: Optional[torch.Tensor] = None    # (77,)           AttentionRefine.alphas
    replace_matrix: Optional[torch.Tensor] = None  # (77,77)         AttentionReplace.mapper
    equalizer: Optional[torch.Tensor] = None       # (77,)           AttentionReweight.equalizer
    blend_alpha: Optional[torch.Tensor] = None     # (2,77)          LocalBlend.alpha_layers
    start_blend: int = 0
    blend_th: float = 0.3


def make_edit_spec(prompts, is_replace_controller, cross_replace_steps, self_replace_steps, blend_word=None,
                   equilizer_params=None, num_steps=None, tokenizer=None) -> EditSpec:
    """p2p/ptp_controller_utils.py:106-133 + p2p/ptp_classes.py:18-42,164-182,229-283."""
    sa = (0.0, self_replace_steps) if isinstance(self_replace_steps, float) else self_replace_steps
    spec = EditSpec(
        num_steps=num_steps, is_replace=bool(is_replace_controller),
        alpha_words=cross_alpha_table(prompts, num_steps, cross_replace_steps, tokenizer),
        self_window=(int(num_steps * sa[0]), int(num_steps * sa[1])))
    if is_replace_controller:
        spec.replace_matrix = replacement_mapper(prompts, tokenizer)
    else:
        spec.mapper, spec.refine_alpha = refinement_mapper(prompts, tokenizer)
    if equilizer_params is not None:
        spec.equalizer = equalizer_row(prompts[1], equilizer_params["words"], equilizer_params["values"], tokenizer)
    if blend_word is not None:
        al = torch.zeros(2, MAX_WORDS)
        for r, (prompt, words) in enumerate(zip(prompts, blend_word)):
            for w in ([words] if isinstance(words, str) else words):
                al[r, torch.as_tensor(word_token_indices(prompt, w, tokenizer))] = 1
        spec.blend_alpha = al
        spec.start_blend = int(0.2 * num_steps)
    return spec


# ----------------------------------------------------------------------------------------------------------
# run-time state + hook
# ----------------------------------------------------------------------------------------------------------
_KEYS = ("down_cross", "mid_cross", "up_cross", "down_self", "mid_self", "up_self")


@dataclass
class P2PState:
    num_att_layers: int = 32
    cur_step: int = 0
    cur_att_layer: int = 0
    blend_calls: int = 0
    step_store: Dict[str, list] = field(default_factory=lambda: {k: [] for k in _KEYS})
    attention_store: Dict[str, list] = field(default_factory=dict)


def _mapped_base(spec: EditSpec, base: torch.Tensor, tar: torch.Tensor) -> torch.Tensor:
    """replace_cross_attention of Refine (ptp_classes.py:259-262) / Replace (:241-243), then the optional
    Reweight wrapper (:279-283).  base, tar: (heads, N, 77)."""
    if spec.is_replace:
        m = torch.einsum("hpw,wn->hpn", base, spec.replace_matrix)
    else:
        m = base[:, :, spec.mapper] * spec.refine_alpha + tar * (1 - spec.refine_alpha)
    if spec.equalizer is not None:
        m = m * spec.equalizer
    return m


def p2p_hook(state: P2PState, spec: EditSpec, probs: torch.Tensor, is_cross: bool, place: str, save_attn: bool):
    """AttentionControl.__call__ (ptp_classes.py:91-108) + AttentionControlEdit.forward (:202-227).
    `probs` is (4*heads, N, M) for the batch [uncond-src, uncond-tar, cond-src, cond-tar]; edited IN PLACE."""
    half = probs.shape[0] // 2
    cond = probs[half:]                                   # view: [src heads..., tar heads...]
    heads = cond.shape[0] // 2
    if cond.shape[1] <= 32 ** 2 and save_attn:            # AttentionStore.forward (:135-141); stores a VIEW
        state.step_store[f"{place}_{'cross' if is_cross else 'self'}"].append(cond)
    lo, hi = spec.self_window
    if is_cross:
        base, tar = cond[:heads], cond[heads:]
        aw = spec.alpha_words[state.cur_step]
        cond[heads:] = _mapped_base(spec, base, tar) * aw + (1 - aw) * tar
    elif lo <= state.cur_step < hi and cond.shape[2] <= 32 ** 2:   # replace_self_attention (:194-200)
        cond[heads:] = cond[:heads]
    if not save_attn:
        return
    state.cur_att_layer += 1
    if state.cur_att_layer == state.num_att_layers:
        state.cur_att_layer = 0
        state.cur_step += 1
        if not state.attention_store:                     # between_steps (:143-150)
            state.attention_store = state.step_store
        else:
            for k in state.attention_store:
                for i in range(len(state.attention_store[k])):
                    state.attention_store[k][i] += state.step_store[k][i]
        state.step_store = {k: [] for k in _KEYS}


def local_blend(state: P2PState, spec: EditSpec, x_t: torch.Tensor) -> torch.Tensor:
    """AttentionControlEdit.step_callback (ptp_classes.py:189-192) -> LocalBlend.__call__/get_mask (:44-72)."""
    if spec.blend_alpha is None:
        return x_t
    state.blend_calls += 1
    if state.blend_calls <= spec.start_blend:
        return x_t
    maps = state.attention_store["down_cross"][2:4] + state.attention_store["up_cross"][:3]
    maps = torch.cat([m.reshape(2, -1, 1, 16, 16, MAX_WORDS) for m in maps], dim=1)
    m = (maps * spec.blend_alpha.reshape(2, 1, 1, 1, 1, MAX_WORDS)).sum(-1).mean(1)
    m = F.max_pool2d(m, (3, 3), (1, 1), padding=(1, 1))
    m = F.interpolate(m, size=x_t.shape[2:])
    m = m / m.max(2, keepdim=True)[0].max(3, keepdim=True)[0]
    m = m.gt(spec.blend_th)
    m = (m[:1] + m).float()
    return x_t[:1] + m * (x_t - x_t[:1])


class OracleP2PProcessor:
    """p2p/ptp_utils.py:31-123 for the SD-1.x Attention configuration (no spatial/group norm, no norm_cross,
    no residual connection, rescale factor 1)."""

    def __init__(self, state: P2PState, spec: EditSpec, place: str):
        self.state, self.spec, self.place = state, spec, place

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 use_controller=True, save_attn=True):
        is_cross = encoder_hidden_states is not None
        ctx = encoder_hidden_states if is_cross else hidden_states
        q = attn.head_to_batch_dim(attn.to_q(hidden_states))
        k = attn.head_to_batch_dim(attn.to_k(ctx))
        v = attn.head_to_batch_dim(attn.to_v(ctx))
        probs = attn.get_attention_scores(q, k, attention_mask)
        if use_controller:
            p2p_hook(self.state, self.spec, probs, is_cross, self.place, save_attn)
        out = attn.batch_to_head_dim(torch.bmm(probs, v))
        return attn.to_out[1](attn.to_out[0](out))


def install_oracle_p2p(unet, state: P2PState, spec: EditSpec) -> None:
    """register_attention_control (ptp_utils.py:277-295)."""
    procs, count = {}, 0
    for name in unet.attn_processors.keys():
        place = "mid" if name.startswith("mid_block") else "up" if name.startswith("up_blocks") else \
            "down" if name.startswith("down_blocks") else None
        if place is None:
            continue
        count += 1
        procs[name] = OracleP2PProcessor(state, spec, place)
    unet.set_attn_processor(procs)
    state.num_att_layers = count
