"""Shared helper: run the oracle port on a golden fixture's inputs."""
import os

import torch

from oracle import h_edit as oh
from oracle import p2p as op
from oracle.pipeline import OraclePipeline
from oracle.sd_unet import UNetConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    # The goldens were generated with 8 intra-op threads (meta["threads"]); oneDNN/MKL partition their fp32 reductions by thread count, so
    # re-running the oracle with a different count perturbs results at the 1e-4 level -- enough to flip a thresholded LocalBlend mask
    # pixel.  Pin the count so that the CPU pins stay bit-exact on any host.
    if not torch.cuda.is_available():
        torch.set_num_threads(int(g["meta"].get("threads", 8)))
    return g


def cfg_from_meta(meta) -> UNetConfig:
    u = meta["unet"]
    return UNetConfig(sample_size=u["sample_size"], block_out_channels=tuple(u["block_out_channels"]),
                      cross_attention_dim=u["cross_attention_dim"], attention_head_dim=u["heads"])


def spec_from_meta(meta, tokenizer) -> op.EditSpec:
    blend = meta["blend"]
    bw = meta["blend_words"]
    K = meta["K"]
    return op.make_edit_spec(
        meta["prompts"], meta["is_replace"], meta["xa"], meta["sa"],
        blend_word=((bw[0],), (bw[1],)) if blend else None,
        equilizer_params={"words": (bw[1],), "values": (1.25 if K > 1 else 2.0,)} if blend else None,
        num_steps=meta["T"], tokenizer=tokenizer)


def run_oracle_on_golden(g, steps=None):
    """Returns (edited, recon, trace) from the oracle port fed with the fixture's xT / zs."""
    meta = g["meta"]
    model = OraclePipeline(cfg_from_meta(meta), seed=0)
    model.scheduler.set_timesteps(meta["T"])
    spec = spec_from_meta(meta, model.tokenizer)
    trace = []
    ed, rc = oh.h_edit_p2p_implicit(
        model.unet, model.scheduler, g["ctx_uncond"], g["ctx_src"], g["ctx_tar"], g["xT"], g["zs"], spec,
        meta["cfg_scales"], eta=meta["eta"], weight_reconstruction=meta["weight_reconstruction"],
        optimization_steps=meta["K"], after_skip_steps=meta["T"], is_ddim_inversion=False, trace=trace)
    return ed, rc, torch.stack(trace), spec
