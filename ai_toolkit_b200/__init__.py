"""B200-native LoRA-training hot path behind ostris/ai-toolkit's plugin API.

Public surface (mirrors the reference):
  LoRASpecialNetwork, LoRAModule, get_network   toolkit/lora_special.py, BaseSDTrainProcess.py:1926-1993
  B200AdamW                                     toolkit/optimizer.py:78-79 (AdamW, eps 1e-6) + clip + EMA, fused
  FluxTransformer2DModel, FluxLoRATrainStep     the frozen DiT and one optimizer step (SDTrainer.hook_train_loop)
  cabi / ops                                    the C ABI (include/b200_lora.h) and tensor-level wrappers

Importing the package never needs a GPU; running an ACTIVE network does (no CPU / eager fallback).
"""
from .lora_special import LoRAModule, LoRASpecialNetwork, get_network  # noqa: F401

__version__ = "0.1.0"
