"""ctypes binding of libhedit_b200.so (C ABI declared in include/hedit_b200.h).

There is NO CPU fallback: if the shared library (built by `__graft_entry__.build()` / `make -C hedit_b200/csrc`)
is missing, or no sm_100a device is visible when an engine is created, this raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhedit_b200.so")


class UNetConfigC(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("sample_size", C.c_int32),
                ("block_out_channels", C.c_int32 * 4), ("layers_per_block", C.c_int32), ("heads", C.c_int32),
                ("cross_attention_dim", C.c_int32), ("norm_groups", C.c_int32), ("ctx_len", C.c_int32)]


class VaeConfigC(C.Structure):
    _fields_ = [("latent_channels", C.c_int32), ("out_channels", C.c_int32), ("block_out_channels", C.c_int32 * 4),
                ("layers_per_block", C.c_int32), ("norm_groups", C.c_int32)]


class TextConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("vocab", "width", "heads", "layers", "ffn", "tokens")]


class ClipConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("resolution", "patch", "width", "heads", "layers")]


class FaceConfigC(C.Structure):
    _fields_ = [("ch", C.c_int32), ("n_levels", C.c_int32), ("ch_mult", C.c_int32 * 8), ("num_res_blocks", C.c_int32), ("attn_resolution", C.c_int32),
                ("image_size", C.c_int32), ("in_channels", C.c_int32), ("out_ch", C.c_int32)]


class FaceStepCoefC(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("t", "tm1", "sqrt_1m_at", "sqrt_at", "sqrt_1m_atm1", "sqrt_atm1", "c2", "noise")]


REWARD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int)


class FaceArgsC(C.Structure):
    _fields_ = [("B", C.c_int32), ("steps", C.c_int32), ("opt_steps", C.c_int32), ("xT", C.c_void_p), ("zs", C.c_void_p), ("coef", C.c_void_p),
                ("weight", C.c_float), ("mask", C.c_void_p), ("use_id", C.c_int32), ("use_lpips", C.c_int32), ("reward", C.c_void_p),
                ("reward_user", C.c_void_p), ("reward_x0", C.c_void_p), ("reward_grad", C.c_void_p), ("edited", C.c_void_p),
                ("n_sample_forwards", C.c_int64), ("n_kernel_launches", C.c_int64)]


class StepCoefC(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("sqrt_1m_at", "sqrt_at", "sqrt_ap", "dir", "noise", "coeff")]


GUIDANCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int)
# hedit_attn_probs_fn(user, tf_index, is_cross, place, probs, batch_heads, n_query, n_key)
ATTN_PROBS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int)
# hedit_attn_editor_fn(user, tf_index, is_cross, place, q, k, v, sim, attn, out, batch_heads, n_query, n_key, d)
ATTN_EDITOR_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_int, C.c_int, C.c_int, C.c_int)


class EditArgsC(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("steps", C.c_int32), ("opt_steps", C.c_int32), ("explicit_form", C.c_int32),
        ("schedule", C.c_int32), ("buffers_on_host", C.c_int32), ("variant", C.c_int32),
        ("xT", C.c_void_p), ("zs", C.c_void_p), ("ctx", C.c_void_p), ("timesteps", C.c_void_p), ("coef", C.c_void_p),
        ("w_src", C.c_float), ("w_src_edit", C.c_float), ("w_tar", C.c_float), ("weight_reconstruction", C.c_float),
        ("use_p2p", C.c_int32),
        ("mapper", C.c_void_p), ("map_w", C.c_void_p), ("map_rows", C.c_int32), ("is_replace", C.c_void_p), ("replace_m", C.c_void_p), ("c_base", C.c_void_p), ("c_tar", C.c_void_p),
        ("self_lo", C.c_int32), ("self_hi", C.c_int32), ("self_max_tokens", C.c_int32),
        ("has_blend", C.c_void_p), ("blend_alpha", C.c_void_p), ("start_blend", C.c_int32), ("blend_th", C.c_float),
        ("blend_rows", C.c_int32), ("blend_th_sub", C.c_float),
        ("masa", C.c_int32), ("masa_layer_mask", C.c_uint32), ("masa_step_on", C.c_void_p), ("mos_pull", C.c_int32),
        ("pnp", C.c_int32), ("pnp_self_mask", C.c_uint32), ("pnp_qk_on", C.c_void_p), ("pnp_feat_on", C.c_void_p),
        ("pre_step", C.c_int32), ("pre_coeff", C.c_float),
        ("guidance", C.c_void_p), ("guidance_user", C.c_void_p), ("guidance_weight", C.c_float), ("x0_coef", C.c_void_p),
        ("guid_x0", C.c_void_p), ("guid_grad", C.c_void_p),
        ("xt_is_pair", C.c_int32), ("ctrl_step0", C.c_int32), ("blend_state", C.c_void_p),
        ("edited", C.c_void_p), ("recon", C.c_void_p), ("trace", C.c_void_p),
        ("n_sample_forwards", C.c_int64), ("n_kernel_launches", C.c_int64),
        ("coef_edit", C.c_void_p),
    ]


# every symbol include/hedit_b200.h declares: (name, restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int, C.c_float
SYMBOLS = {
    "hedit_last_error": (C.c_char_p, []),
    "hedit_device_count": (_I, []),
    "hedit_engine_create": (_P, [C.POINTER(UNetConfigC), _I, _I, _I]),
    "hedit_engine_destroy": (None, [_P]),
    "hedit_engine_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_engine_finalize": (_I, [_P]),
    "hedit_engine_flops_per_sample": (C.c_double, [_P]),
    "hedit_engine_blend_state_elems": (_I, [_P, _I]),
    "hedit_operand_dtype": (C.c_char_p, []),
    "hedit_engine_tensor_count": (_I, [_P]),
    "hedit_engine_tensor_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(C.c_int64)]),
    "hedit_engine_profile_forward": (_I, [_P, _I, _I, C.c_char_p, _I]),
    "hedit_engine_set_graph_replay": (_I, [_P, _I]),
    "hedit_engine_set_prefix_dedup": (_I, [_P, _I]),
    "hedit_engine_set_splitk": (_I, [_P, _I]),
    "hedit_abi_sizeof": (_I, [C.c_char_p]),
    "hedit_unet_forward": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "hedit_unet_forward_indexed": (_I, [_P, _P, _P, _P, _I, _P, _I, _P, _P]),
    "hedit_unet_forward_compat": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _P]),
    "hedit_unet_forward_editor": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _P]),
    "hedit_edit_p2p": (_I, [_P, C.POINTER(EditArgsC), _P]),
    "hedit_vae_create": (_P, [C.POINTER(VaeConfigC), _I]),
    "hedit_vae_destroy": (None, [_P]),
    "hedit_vae_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_vae_finalize": (_I, [_P]),
    "hedit_vae_tensor_count": (_I, [_P]),
    "hedit_vae_tensor_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(C.c_int64)]),
    "hedit_vae_decode": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "hedit_vae_decode_backward": (_I, [_P, _P, _P, _P]),
    "hedit_vae_last_flops": (C.c_double, [_P]),
    "hedit_text_create": (_P, [C.POINTER(TextConfigC), _I]),
    "hedit_text_destroy": (None, [_P]),
    "hedit_text_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_text_finalize": (_I, [_P]),
    "hedit_text_encode": (_I, [_P, _P, _I, _P, _P]),
    "hedit_clip_create": (_P, [C.POINTER(ClipConfigC), _I]),
    "hedit_clip_destroy": (None, [_P]),
    "hedit_clip_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_clip_finalize": (_I, [_P]),
    "hedit_clip_set_reference": (_I, [_P, _P, _P]),
    "hedit_clip_gram_loss": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "hedit_clip_gram_backward": (_I, [_P, _P, _P]),
    "hedit_vae_enc_create": (_P, [C.POINTER(VaeConfigC), _I]),
    "hedit_vae_enc_destroy": (None, [_P]),
    "hedit_vae_enc_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_vae_enc_finalize": (_I, [_P]),
    "hedit_vae_encode": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "hedit_face_create": (_P, [C.POINTER(FaceConfigC), _I]),
    "hedit_face_destroy": (None, [_P]),
    "hedit_face_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_face_finalize": (_I, [_P]),
    "hedit_face_tensor_count": (_I, [_P]),
    "hedit_face_tensor_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(C.c_int64)]),
    "hedit_face_unet_forward": (_I, [_P, _P, _P, _I, _P, _P]),
    "hedit_face_last_flops": (C.c_double, [_P]),
    "hedit_face_edit": (_I, [_P, C.POINTER(FaceArgsC), _P]),
    "hedit_arcface_create": (_P, [_I]),
    "hedit_arcface_destroy": (None, [_P]),
    "hedit_arcface_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_arcface_finalize": (_I, [_P]),
    "hedit_arcface_features": (_I, [_P, _P, _I, _P, _P]),
    "hedit_arcface_set_reference": (_I, [_P, _P, _P]),
    "hedit_arcface_loss_grad": (_I, [_P, _P, _I, _P, _P, _P]),
    "hedit_arcface_last_flops": (C.c_double, [_P]),
    "hedit_lpips_create": (_P, [_I]),
    "hedit_lpips_destroy": (None, [_P]),
    "hedit_lpips_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hedit_lpips_finalize": (_I, [_P]),
    "hedit_lpips_set_source": (_I, [_P, _P, _I, _I, _P]),
    "hedit_lpips_loss_grad": (_I, [_P, _P, _I, _P, _P, _P]),
    "hedit_lpips_last_flops": (C.c_double, [_P]),
    "hedit_op_linear": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "hedit_op_linear_geglu": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "hedit_op_conv3x3": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "hedit_op_self_attention": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "hedit_op_cross_attention_p2p": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "hedit_op_group_norm": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "hedit_op_layer_norm": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
}

_lib = None


def load():
    """Load the shared library (once) and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"hedit_b200: native library not found at {LIB_PATH}. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C hedit_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def operand_torch_dtype():
    import torch
    return torch.float16 if load().hedit_operand_dtype() == b"fp16" else torch.bfloat16


def last_error() -> str:
    return load().hedit_last_error().decode()


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise RuntimeError(f"hedit_b200: {what} failed ({rc}): {last_error()}")
    return rc
