"""LDM <-> diffusers LoRA key maps (`toolkit/network_mixins.py:525-579`).

The reference ships its key maps as JSON data under `toolkit/keymaps/`.  They are data of the host
installation, not code of this package: when this package runs as an ai-toolkit extension the directory is
found next to the host's `toolkit` package, otherwise `AITK_KEYMAPS_ROOT` names it.  The default lookup
(`stable_diffusion_locon_<tail>.json`) finds no file in the reference tree either, so keys are saved as they are
(kohya names for UNet models, peft names for transformers)."""
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
    if use_weight_mapping:
        print("[b200] weight-mapped LDM key conversion (ssd / vega / forced) is not implemented; keys are kept as they are")
        return None
    name = f"stable_diffusion_locon_{tail}.json"
    for root in _roots():
        path = os.path.join(root, name)
        if os.path.exists(path):
            with open(path, "r") as f:
                return json.load(f)["ldm_diffusers_keymap"]
    return None
