"""hedit_b200 -- B200-native implementation of h-Edit's reverse-time bridge sampling loop (nktoan/h-edit hot path).

Layout: csrc/ (hand-written sm_100a kernels + C ABI), _lib.py (ctypes binding), engine.py (engine handle),
p2p.py (host-side Prompt-to-Prompt set-up -> device edit plan), schedule.py (per-step scalar tables),
samplers.py (reference-compatible h_Edit_* callables)."""
from .p2p import (EditController, LocalBlend, compile_edit_plan, get_equalizer, get_refinement_mapper,  # noqa: F401
                  get_replacement_mapper, get_time_words_attention_alpha, get_word_inds, make_controller,
                  register_attention_control)
from .schedule import DDIMTables, skip_pre_coeff, step_tables, x0_tables  # noqa: F401
from .p2p import clear_setup_cache  # noqa: F401
from .tokenizer import WordTokenizer  # noqa: F401
from .engine import UNetEngine, unet_config_of  # noqa: F401
from .samplers import (encode_prompts, invalidate_engines, HEditStepper, h_edit_step, MutualSelfAttentionControl, encode_text, get_engine, h_Edit_masactrl_explicit, h_Edit_masactrl_implicit, h_Edit_p2p_explicit,  # noqa: F401
                       h_Edit_p2p_implicit, h_Edit_R_explicit, h_Edit_R_implicit, h_edit_p2p_batch, regiter_attention_editor_diffusers,
                       h_Edit_PnP_implicit, ef_or_pnp_inv_w_p2p, ef_wo_p2p, ef_or_pnp_inv_w_masactrl, ef_or_pnp_inv_w_pnp, negative_prompt_pnp, nmg_p2p, nmg_pnp, nulltext_pnp, pnp_self_mask, pnp_step_flags, register_attention_control_efficient, register_conv_control_efficient, register_time)

from . import style  # noqa: F401,E402
from .vae import VaeDecoderEngine, VaeEncoderEngine, vae_config_of  # noqa: F401,E402
from .clip_gram import ClipGramEngine  # noqa: F401,E402
from .text_encoder import TextEncoderEngine  # noqa: F401,E402
from . import face  # noqa: F401,E402
from .face import FaceUNetEngine  # noqa: F401,E402
from . import compat  # noqa: F401,E402
from .compat import CompatUNet, controller_kind, h_edit_p2p_implicit_compat, register_attention_control_compat  # noqa: F401,E402
from .pipeline import SyntheticPipeline, random_text_engine  # noqa: F401,E402
from .inversion import ddim_inversion, inversion_forward_process_ddpm, sample_xts_from_x0  # noqa: F401,E402

__all__ = ["inversion_forward_process_ddpm", "ddim_inversion", "UNetEngine", "unet_config_of", "make_controller", "register_attention_control", "compile_edit_plan",
           "h_Edit_p2p_implicit", "h_Edit_p2p_explicit", "h_Edit_R_implicit", "h_Edit_R_explicit", "h_Edit_masactrl_implicit", "h_Edit_masactrl_explicit", "h_Edit_PnP_implicit", "register_attention_control_efficient", "register_conv_control_efficient", "register_time",
           "MutualSelfAttentionControl", "regiter_attention_editor_diffusers", "h_edit_p2p_batch", "h_edit_step", "HEditStepper", "encode_text", "step_tables"]
