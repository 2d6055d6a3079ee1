"""LDM <-> diffusers LoRA key maps (`toolkit/network_mixins.py:525-579`).

The reference ships its key maps as JSON data under `toolkit/keymaps/`.  They are data of the host
installation, not code of this package: when this package runs as an ai-toolkit extension the directory is
found next to the host's `toolkit` package, otherwise `AITK_KEYMAPS_ROOT` names it.  The default lookup
(`stable_diffusion_locon_<tail>.json`) finds no file in the reference tree either, so keys are saved as they are
(kohya names for UNet models, peft names for transformers); the weight-mapped conversion (ssd / vega models, or the forced
retry of `load_weights` for SD1 files with unknown keys) derives the LoRA key map from `stable_diffusion_<tail>.json`."""
from __future__ import annotations

import json
import os


def _roots():
    roots = []
    env = os.environ.get("AITK_KEYMAPS_ROOT")
    if env:
        roots.append(env)
    try:
        import toolkit  # the host ai-toolkit installation, when present

        roots.append(os.path.join(os.path.dirname(toolkit.__file__), "keymaps"))
    except Exception:
        pass
    return roots


def lora_keymap_from_model_keymap(model_keymap):
    """`get_lora_keymap_from_model_keymap` (toolkit/saving.py:279-330): a full-model LDM -> diffusers key map turned into
    the kohya LoRA key map (`lora_unet_*` / `lora_te*` with underscores; one entry per adapter tensor + alpha)."""
    from collections import OrderedDict

    lora_keymap = OrderedDict()
    dual = any(k.startswith("conditioner.embedders.1") for k in model_keymap)
    for key, value in model_keymap.items():
        if key.endswith("bias"):
            continue
        if key.endswith(".weight"):
            key = key[:-7]
        if value.endswith(".weight"):
            value = value[:-7]
        key = key.replace("model.diffusion_model", "lora_unet")
        if value.startswith("unet"):
            value = f"lora_{value}"
        if dual:
            key = key.replace("conditioner.embedders.0", "lora_te1").replace("conditioner.embedders.1", "lora_te2")
            if value.startswith("te0") or value.startswith("te1"):
                value = f"lora_{value}"
            # (the reference's two `value.replace(...)` calls at :311-312 discard their result: no renaming happens)
        key = key.replace("cond_stage_model.transformer", "lora_te")
        if value.startswith("te_"):
            value = f"lora_{value}"
        key, value = key.replace(".", "_"), value.replace(".", "_")
        for tail in ("lora_down.weight", "lora_down.bias", "lora_up.weight", "lora_up.bias", "alpha"):
            lora_keymap[f"{key}.{tail}"] = f"{value}.{tail}"
    return lora_keymap


def _find(name):
    for root in _roots():
        path = os.path.join(root, name)
        if os.path.exists(path):
            with open(path, "r") as f:
                return json.load(f)["ldm_diffusers_keymap"]
    return None


def load_keymap(network, force_weight_mapping=False):
    use_weight_mapping = False
    if network.is_ssd:
        tail, use_weight_mapping = "ssd", True
    elif network.is_vega:
        tail, use_weight_mapping = "vega", True
    elif network.is_sdxl:
        tail = "sdxl"
    elif network.is_v2:
        tail = "sd2"
    else:
        tail = "sd1"
    if force_weight_mapping:
        use_weight_mapping = True
    if use_weight_mapping:  # ssd / vega / forced: derive the LoRA key map from the full-model key map (:564-566)
        keymap = _find(f"stable_diffusion_{tail}.json")
        return lora_keymap_from_model_keymap(keymap) if keymap is not None else None
    return _find(f"stable_diffusion_locon_{tail}.json")
