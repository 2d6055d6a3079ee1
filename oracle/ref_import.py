"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference (`/root/reference`, ostris/ai-toolkit
@ 27a03a9) in this container so that its own `LoRASpecialNetwork` / `LoRAModule` /
`ToolkitModuleMixin.forward` can generate golden vectors.

The reference imports third-party packages that are not installed here (diffusers, optimum.quanto,
torchao, ...).  They are only touched for type names and `isinstance(x, QTensor)` checks on this path,
so a `sys.meta_path` finder fabricates empty stub modules for those roots.  Nothing under
`ai_toolkit_b200/` imports this file; only `oracle/make_golden.py` (run here, outputs committed under
`tests/golden/`) and CPU tests that are skipped when `/root/reference` is absent (the GPU box).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AITK_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = {
    "diffusers", "optimum", "torchao", "torchaudio", "av", "lycoris", "peft", "accelerate", "bitsandbytes",
    "prodigyopt", "oyaml", "flatten_json", "omegaconf", "kornia", "albumentations", "lpips", "open_clip", "timm",
    "pytorch_wavelets", "torchcodec", "librosa", "mutagen", "controlnet_aux", "cv2", "k_diffusion", "dctorch",
    "wandb", "tensorboard", "gguf", "torch_xla",
}


class _Any:
    """Instance returned by any call on a stub class: every attribute / call yields another `_Any`."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


class _StubMeta(type):
    """Stub classes answer class-level attribute access (`logging.get_logger(...)`, `Scheduler.from_config(...)`) with a
    permissive callable, stay subclassable (`class CustomLCMScheduler(LCMScheduler)`) and remain real classes for the
    `isinstance(x, QTensor)` checks on the LoRA path (False for every real tensor)."""

    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: _Any()


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        cls = _StubMeta(name, (), {"__module__": self.__name__, "__init__": lambda s, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        root = name.split(".")[0]
        if root in _STUB_ROOTS:
            try:  # a real installation wins
                for f in sys.meta_path:
                    if f is self:
                        continue
                    spec = f.find_spec(name, path, target) if hasattr(f, "find_spec") else None
                    if spec is not None:
                        return spec
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "toolkit"))


def install():
    """Make `import toolkit.lora_special` work.  Import torch/transformers first (Appendix B)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch  # noqa: F401
    import transformers  # noqa: F401
    # transformers resolves its classes lazily and probes optional packages with find_spec(): resolve what
    # the reference needs BEFORE the stub finder can make absent packages look installed.
    from transformers import (CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer, T5EncoderModel,  # noqa: F401
                              T5Tokenizer, UMT5EncoderModel)
    sys.meta_path.append(_StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def reference_sd_trainer():
    """The reference's `SDTrainer` class, unmodified (extensions_built_in/sd_trainer/SDTrainer.py).  It imports with the
    stubs; instances are made with `object.__new__` + the attributes a method reads (its __init__ needs a real job)."""
    install()
    from extensions_built_in.sd_trainer.SDTrainer import SDTrainer  # type: ignore

    return SDTrainer


def reference_lora():
    """Return the reference's (LoRASpecialNetwork, LoRAModule) classes, unmodified."""
    install()
    from toolkit.lora_special import LoRAModule, LoRASpecialNetwork  # type: ignore

    return LoRASpecialNetwork, LoRAModule
