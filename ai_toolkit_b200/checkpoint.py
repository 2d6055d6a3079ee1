"""Save / resume of a LoRA run in the reference's on-disk conventions (host logic only).

What the reference's trainer does around `network.save_weights` / `load_weights`, restated as functions because the
trainer class itself (`BaseSDTrainProcess`) needs accelerate + diffusers to exist:
  * `save_checkpoint`   = `BaseSDTrainProcess.save` :505-721 for a LoRA run: `{name}_{step:09d}.safetensors` (no step
    suffix for the final save) written with multiplier 1.0, metadata `training_info` {step, epoch} +
    `ss_base_model_version` + `ss_output_name` (`update_training_metadata` :388-409) flattened by
    `get_meta_for_safetensors`, then `optimizer.pt` = the optimizer's `state_dict()` (:702-714), then `clean_up_saves`.
  * `clean_up_saves`    = :418-493: keep the newest `max_step_saves_to_keep` step files of each kind, by ctime.
  * `get_latest_save_path` = :822-865: newest `{name}*{post}` file or folder, minus the `_LoRA/_refiner/_t2i/_cn`
    false positives, falling back to `pretrained_lora_path`.
  * `load_training_state_from_metadata` = :867-889: `(step, epoch)` from the file's `training_info`.
  * `load_metadata_from_safetensors` / `parse_metadata_from_safetensors` = toolkit/metadata.py:70-88.
`B200AdamW.torch_state_dict()` is the `torch.optim.AdamW` layout, so `optimizer.pt` written here loads into the
reference's optimizer and vice versa.
"""
from __future__ import annotations

import copy
import glob
import json
import os
import shutil
from collections import OrderedDict
from typing import Optional

import torch

from .metadata import get_meta_for_safetensors


def parse_metadata_from_safetensors(meta) -> OrderedDict:
    parsed = OrderedDict()
    for key, value in (meta or {}).items():
        try:
            parsed[key] = json.loads(value)
        except json.decoder.JSONDecodeError:
            parsed[key] = value
    return parsed


def load_metadata_from_safetensors(file_path: str) -> OrderedDict:
    from safetensors import safe_open

    try:
        with safe_open(file_path, framework="pt") as f:
            return parse_metadata_from_safetensors(f.metadata())
    except Exception as e:  # the reference swallows the error and returns an empty dict (:86-88)
        print(f"Error loading metadata from {file_path}: {e}")
        return OrderedDict()


def training_metadata(name: str, step: int, epoch: int = 0, base_model_version: str = "flux.1",
                      trigger_word: Optional[str] = None, extra: Optional[dict] = None) -> OrderedDict:
    meta = OrderedDict(training_info=OrderedDict(step=step, epoch=epoch))
    meta["ss_base_model_version"] = base_model_version
    meta["ss_output_name"] = name
    if trigger_word is not None:  # "just so auto1111 will pick it up" (:401-407)
        meta["ss_tag_frequency"] = {f"1_{trigger_word}": {f"{trigger_word}": 1}}
    if extra:
        meta.update(extra)
    return meta


def clean_up_saves(save_root: str, name: str, max_step_saves_to_keep: int):
    latest_item = None
    if not os.path.exists(save_root):
        return latest_item
    items = glob.glob(os.path.join(save_root, f"{name}_*"))
    safetensors_files = sorted((f for f in items if f.endswith(".safetensors")), key=os.path.getctime)
    pt_files = sorted((f for f in items if f.endswith(".pt")), key=os.path.getctime)
    directories = sorted((d for d in items if os.path.isdir(d) and not d.endswith(".safetensors")), key=os.path.getctime)
    critic_items = sorted(glob.glob(os.path.join(save_root, f"CRITIC_{name}_*")), key=os.path.getctime)
    combined = sorted(safetensors_files + directories + pt_files, key=os.path.getctime)
    keep = max_step_saves_to_keep
    to_remove = []
    for group in (safetensors_files, pt_files, directories, critic_items):
        to_remove += group[:-keep] if group else []
    for item in dict.fromkeys(to_remove):
        if os.path.isdir(item):
            shutil.rmtree(item)
        else:
            os.remove(item)
        yaml_file = os.path.splitext(item)[0] + ".yaml"
        if os.path.exists(yaml_file):
            os.remove(yaml_file)
    if combined:
        latest_item = combined[-1]
    return latest_item


def save_checkpoint(network, optimizer, save_root: str, name: str, step: Optional[int] = None, epoch: int = 0,
                    dtype=torch.float16, meta: Optional[dict] = None, max_step_saves_to_keep: int = 4,
                    named_lora: bool = False, base_model_version: str = "flux.1") -> str:
    os.makedirs(save_root, exist_ok=True)
    step_num = "" if step is None else f"_{str(step).zfill(9)}"
    save_meta = copy.deepcopy(meta) if meta is not None else OrderedDict()
    # `update_training_metadata` does `self.meta.update(...)` (BaseSDTrainProcess.py:388-409): the CURRENT step / epoch /
    # ss_* keys overwrite whatever a loaded checkpoint's metadata carried
    save_meta.update(training_metadata(name, 0 if step is None else step, epoch, base_model_version))
    save_meta = get_meta_for_safetensors(save_meta, name)
    lora_name = name + ("_LoRA" if named_lora else "")
    file_path = os.path.join(save_root, f"{lora_name}{step_num}.safetensors")
    prev_multiplier = network.multiplier
    network.multiplier = 1.0
    try:
        network.save_weights(file_path, dtype=dtype, metadata=save_meta)
    finally:
        network.multiplier = prev_multiplier
    if optimizer is not None:
        sd = optimizer.torch_state_dict() if hasattr(optimizer, "torch_state_dict") else optimizer.state_dict()
        torch.save(sd, os.path.join(save_root, "optimizer.pt"))
    clean_up_saves(save_root, name, max_step_saves_to_keep)
    return file_path


def get_latest_save_path(save_root: str, name: str, post: str = "", pretrained_lora_path: Optional[str] = None):
    latest_path = None
    if os.path.exists(save_root):
        paths = []
        for pattern in (f"{name}*{post}.safetensors", f"{name}*{post}.pt", f"{name}*{post}"):
            paths.extend(glob.glob(os.path.join(save_root, pattern)))
        paths = [p for p in paths if os.path.exists(p)]
        for marker in ("_LoRA", "_refiner", "_t2i", "_cn"):
            if marker not in name:
                paths = [p for p in paths if marker not in p]
        if paths:
            latest_path = max(paths, key=os.path.getctime)
    if latest_path is None and pretrained_lora_path is not None and os.path.exists(pretrained_lora_path):
        latest_path = pretrained_lora_path
    return latest_path


def load_training_state_from_metadata(path: str, pretrained_lora_path: Optional[str] = None):
    """-> (step, epoch) or None (no training_info, or `path` is the pretrained LoRA the run merely starts from)."""
    if path is None or path == pretrained_lora_path:
        return None
    if os.path.isdir(path):
        meta_path = os.path.join(path, "aitk_meta.yaml")
        if not os.path.exists(meta_path):
            return None
        import yaml

        with open(meta_path, "r") as f:
            meta = yaml.load(f, Loader=yaml.FullLoader)
    else:
        meta = load_metadata_from_safetensors(path)
    if meta and "training_info" in meta and "step" in meta["training_info"]:
        info = meta["training_info"]
        return int(info["step"]), int(info.get("epoch", 0))
    return None


def resume(network, optimizer, save_root: str, name: str, pretrained_lora_path: Optional[str] = None):
    """The load side of `BaseSDTrainProcess.run`: network weights + step/epoch from the newest save (:2046-2060), then
    `optimizer.pt` if it exists (:2189-2222) unless loading changed the rank (`network.did_change_weights`: the state
    would not match); the learning rates of the CURRENT config win over the saved ones.  Returns (path, step, epoch)."""
    path = get_latest_save_path(save_root, name, pretrained_lora_path=pretrained_lora_path)
    step, epoch = 0, 0
    if path is not None:
        network.load_weights(path)
        if hasattr(network, "mark_params_changed"):
            network.mark_params_changed()
        # the reference builds the EMA after the weights are loaded (load_weights :2053, setup_ema :2229): the shadow
        # starts from the LOADED parameters, not from the fresh init the optimizer was constructed on
        if optimizer is not None and hasattr(optimizer, "reset_ema"):
            optimizer.reset_ema()
        state = load_training_state_from_metadata(path, pretrained_lora_path)
        if state is not None:
            step, epoch = state
    opt_path = os.path.join(save_root, "optimizer.pt")
    if optimizer is not None and os.path.exists(opt_path) and not getattr(network, "did_change_weights", False):
        previous_lrs = [g["lr"] for g in optimizer.param_groups]
        try:
            sd = torch.load(opt_path, weights_only=True)
            if isinstance(sd, dict) and sd.get("b200_flat"):  # written through B200AdamW.state_dict() (e.g. by the
                optimizer.load_state_dict(sd)                 # reference trainer's own save(), which calls state_dict())
            elif hasattr(optimizer, "load_torch_state_dict"):
                optimizer.load_torch_state_dict(sd)
            else:
                optimizer.load_state_dict(sd)
        except Exception as e:  # the reference logs and continues with a fresh optimizer state
            print(f"Failed to load optimizer state from {opt_path}: {e}")
        for g, lr in zip(optimizer.param_groups, previous_lrs):
            g["lr"] = lr
        if hasattr(optimizer, "sync_hyper"):
            optimizer._hyper_host = None
            optimizer.sync_hyper()
    return path, step, epoch
