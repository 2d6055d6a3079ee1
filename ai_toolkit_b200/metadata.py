"""Safetensors metadata helpers with the reference's on-disk conventions (`toolkit/metadata.py:13-46`,
hash definitions `toolkit/train_tools.py:162-185`): flat str->str metadata, `software`, `format: "pt"`,
`sshs_model_hash` (sha256 of the tensor payload) and `sshs_legacy_hash` (sha256 of bytes
[0x100000, 0x110000) of the file, first 8 hex digits)."""
from __future__ import annotations

import hashlib
import json
from collections import OrderedDict
from io import BytesIO

SOFTWARE_META = OrderedDict(name="ai-toolkit-b200", repo="https://github.com/ostris/ai-toolkit", version="b200-r1")


def get_meta_for_safetensors(meta, name=None, add_software_info=True) -> OrderedDict:
    text = json.dumps(meta)
    if name is not None:
        text = text.replace("[name]", name)
    save_meta = json.loads(text, object_pairs_hook=OrderedDict)
    if add_software_info:
        save_meta["software"] = SOFTWARE_META
    for key, value in save_meta.items():
        if not isinstance(value, str):
            save_meta[key] = json.dumps(value)
    save_meta["format"] = "pt"
    return save_meta


def addnet_hash_safetensors(b) -> str:
    h = hashlib.sha256()
    b.seek(0)
    n = int.from_bytes(b.read(8), "little")
    b.seek(n + 8)
    for chunk in iter(lambda: b.read(1024 * 1024), b""):
        h.update(chunk)
    return h.hexdigest()


def addnet_hash_legacy(b) -> str:
    h = hashlib.sha256()
    b.seek(0x100000)
    h.update(b.read(0x10000))
    return h.hexdigest()[0:8]


def add_model_hash_to_meta(state_dict, meta) -> OrderedDict:
    import safetensors.torch

    metadata = {k: v for k, v in meta.items() if k.startswith("ss_")}
    blob = safetensors.torch.save({k: v.contiguous() for k, v in state_dict.items()}, metadata)
    b = BytesIO(blob)
    meta["sshs_model_hash"] = addnet_hash_safetensors(b)
    meta["sshs_legacy_hash"] = addnet_hash_legacy(b)
    return meta
