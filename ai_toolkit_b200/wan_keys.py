"""Wan2.1 LoRA key conversion between diffusers names (what the network holds) and the original Wan names (what is saved):
`Wan21.convert_lora_weights_before_save / _before_load` (toolkit/models/wan21/wan21.py:726-730) ->
toolkit/models/wan21/wan_lora_convert.py:1-75.  Pure string rewriting; pinned against the live reference functions in
tests/test_wan.py.  `WanLoRABaseModel` is the duck-typed `base_model` the network consults (load_lora / convert hooks,
`get_transformer_block_names`)."""
from __future__ import annotations


def convert_to_diffusers(state_dict):
    out = {}
    for key in state_dict:
        new_key = key
        if key.startswith("diffusion_model."):
            new_key = key.replace("diffusion_model.", "transformer.")
        if "self_attn" in new_key:
            new_key = new_key.replace("self_attn", "attn1")
        elif "cross_attn" in new_key:
            new_key = new_key.replace("cross_attn", "attn2")
        parts = new_key.split(".")
        for i, part in enumerate(parts):
            if part in ("q", "k", "v"):
                parts[i] = f"to_{part}"
            elif part == "k_img":
                parts[i] = "add_k_proj"
            elif part == "v_img":
                parts[i] = "add_v_proj"
            elif part == "o":
                parts[i] = "to_out.0"
        new_key = ".".join(parts)
        if "ffn.0" in new_key:
            new_key = new_key.replace("ffn.0", "ffn.net.0.proj")
        elif "ffn.2" in new_key:
            new_key = new_key.replace("ffn.2", "ffn.net.2")
        out[new_key] = state_dict[key]
    return out


def convert_to_original(state_dict):
    out = {}
    for key in state_dict:
        new_key = key
        if key.startswith("transformer."):
            new_key = key.replace("transformer.", "diffusion_model.")
        if "attn1" in new_key:
            new_key = new_key.replace("attn1", "self_attn")
        elif "attn2" in new_key:
            new_key = new_key.replace("attn2", "cross_attn")
        if "to_out.0" in new_key:
            new_key = new_key.replace("to_out.0", "o")
        elif "to_q" in new_key:
            new_key = new_key.replace("to_q", "q")
        elif "to_k" in new_key:
            new_key = new_key.replace("to_k", "k")
        elif "to_v" in new_key:
            new_key = new_key.replace("to_v", "v")
        elif "add_k_proj" in new_key:
            new_key = new_key.replace("add_k_proj", "k_img")
        elif "add_v_proj" in new_key:
            new_key = new_key.replace("add_v_proj", "v_img")
        if "ffn.net.0.proj" in new_key:
            new_key = new_key.replace("ffn.net.0.proj", "ffn.0")
        elif "ffn.net.2" in new_key:
            new_key = new_key.replace("ffn.net.2", "ffn.2")
        out[new_key] = state_dict[key]
    return out


class WanLoRABaseModel:
    """The slice of `Wan21(BaseModel)` that `LoRASpecialNetwork` touches through `base_model`: block names for
    `transformer_only` (toolkit/models/base_model.py `get_transformer_block_names`), the target class list (wan21.py:330) and
    the two key-conversion hooks (:726-730)."""

    arch = "wan21"
    is_transformer = True
    use_old_lokr_format = False  # read by LoRASpecialNetwork.__init__ (toolkit/lora_special.py:419)
    target_lora_modules = ["WanTransformer3DModel"]

    def get_transformer_block_names(self):
        return ["blocks"]

    def convert_lora_weights_before_save(self, state_dict):
        return convert_to_original(state_dict)

    def convert_lora_weights_before_load(self, state_dict):
        return convert_to_diffusers(state_dict)
