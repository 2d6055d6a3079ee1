"""B200-native `LoRASpecialNetwork` / `LoRAModule` behind the reference's plugin API.

Mirrors the interface of ostris/ai-toolkit's `toolkit/lora_special.py:46-135` (LoRAModule),
`:276-775` (LoRASpecialNetwork), `toolkit/network_mixins.py:274-348` (module forward),
`:491-932` (network mixin: multiplier, `with network:`, force_to, merge_in/out, get_state_dict /
save_weights / load_weights) and `toolkit/kohya_lora.py:952-965, 1030-1074` (apply_to,
prepare_optimizer_params): same constructor keywords, same `lora_name`s, same state-dict keys and
`.safetensors` layout, so a config that names this network trains and saves exactly as before.

What is different is where the arithmetic runs:

* every `lora_down.weight` / `lora_up.weight` is a VIEW into one flat fp32 buffer (`flat_params`) and its
  `.grad` a view into `flat_grads`, so grad-norm / clip / AdamW / all-reduce are single launches over a
  contiguous buffer (`B200AdamW`), and bf16 padded operand copies for the tensor cores are refreshed by one
  `b200_repack_lora` launch;
* `LoRAModule.forward` is one fused tcgen05 GEMM (frozen base + rank-r update in the same TMEM tile,
  `b200_gemm_bf16`) plus one rank-side GEMM, and its backward emits dX, dA, dB without forming dW.

There is no eager fallback: on a device without the sm_100a library an ACTIVE network raises
`cabi.B200Error`.  Construction, naming, state-dict, save/load and merge are host logic and run anywhere.
"""
from __future__ import annotations

import math
import os
import weakref
from collections import OrderedDict
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn

from . import cabi

LINEAR_MODULES = ["Linear", "LoRACompatibleLinear", "QLinear", "OstrisLinear"]  # lora_special.py:29-35
CONV_MODULES = ["Conv2d", "LoRACompatibleConv", "QConv2d"]  # lora_special.py:36-40
RANK_PAD = 64
KOHYA_UNET_TARGET_REPLACE_MODULE = ["Transformer2DModel"]  # kohya_lora.py:750
KOHYA_UNET_TARGET_REPLACE_MODULE_CONV2D_3X3 = ["ResnetBlock2D", "Downsample2D", "Upsample2D"]  # :751


class LoRAModule(nn.Module):
    """Adapter for one `nn.Linear`; replaces the Linear's `forward` (lora_special.py:46-135)."""

    def __init__(self, lora_name, org_module: nn.Module, multiplier=1.0, lora_dim=4, alpha=1, dropout=None,
                 rank_dropout=None, module_dropout=None, network: "LoRASpecialNetwork" = None, use_bias: bool = False,
                 **kwargs):
        super().__init__()
        self.can_merge_in = True
        self.network_ref = weakref.ref(network) if network is not None else None
        self.lora_name = lora_name
        self.orig_module_ref = weakref.ref(org_module)
        cls = org_module.__class__.__name__
        self.lora_dim = int(lora_dim)
        if self.lora_dim > RANK_PAD:
            raise NotImplementedError(f"rank {self.lora_dim} > {RANK_PAD}")
        if cls in CONV_MODULES:
            # lora_special.py:95-104: down = Conv2d(in, r, k, stride, padding), up = Conv2d(r, out, 1x1).  Here both (and
            # the frozen conv) are ONE fused GEMM over im2col rows [B Ho Wo, in*kh*kw]; `in_dim` is that contraction length
            if getattr(org_module, "groups", 1) != 1 or tuple(getattr(org_module, "dilation", (1, 1))) != (1, 1):
                raise NotImplementedError(f"{lora_name}: grouped / dilated Conv2d adapters are not implemented")
            if isinstance(org_module.padding, str):
                raise NotImplementedError(f"{lora_name}: string padding modes are not implemented")
            self.kernel_size = tuple(org_module.kernel_size)
            self.stride, self.padding = tuple(org_module.stride), tuple(org_module.padding)
            self.in_channels = org_module.in_channels
            in_dim, out_dim = org_module.in_channels * self.kernel_size[0] * self.kernel_size[1], org_module.out_channels
            if in_dim % 8 != 0 or out_dim % 8 != 0:
                raise NotImplementedError(f"{lora_name}: conv adapter needs in*kh*kw ({in_dim}) and out ({out_dim}) % 8 == 0")
            self.is_conv = True
            self.lora_down = nn.Conv2d(self.in_channels, self.lora_dim, self.kernel_size, self.stride, self.padding, bias=False)
            self.lora_up = nn.Conv2d(self.lora_dim, out_dim, (1, 1), (1, 1), bias=False)
        else:
            in_dim, out_dim = org_module.in_features, org_module.out_features
            self.is_conv = False
            self.lora_down = nn.Linear(in_dim, self.lora_dim, bias=False)
            self.lora_up = nn.Linear(self.lora_dim, out_dim, bias=False)
        if org_module.bias is None:
            use_bias = False
        if use_bias:
            raise NotImplementedError("use_bias (LoRM) adapters are not implemented")
        for nm, pv in (("dropout", dropout), ("rank_dropout", rank_dropout), ("module_dropout", module_dropout)):
            if pv is not None and not isinstance(pv, (int, float)):
                raise NotImplementedError(f"{nm} must be a float or None")
        self.in_dim, self.out_dim = in_dim, out_dim
        self.full_rank = False
        if isinstance(alpha, torch.Tensor):
            alpha = float(alpha.detach().float().item())
        alpha = self.lora_dim if alpha is None or alpha == 0 else alpha
        self.scale = float(alpha) / self.lora_dim
        self.register_buffer("_runtime_scale", torch.tensor(self.scale, dtype=torch.float32), persistent=False)
        self.register_buffer("alpha", torch.tensor(alpha))
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))  # lora_special.py:120
        nn.init.zeros_(self.lora_up.weight)  # :122
        self.multiplier = multiplier
        self.org_module = [org_module]
        self.dropout = dropout
        self.rank_dropout = rank_dropout
        self.module_dropout = module_dropout
        self.is_checkpointing = False
        # filled by LoRASpecialNetwork._flatten(): bf16 operand copies and their parameter versions
        self.a_pack: Optional[torch.Tensor] = None  # [64, in]  rows >= r are zero
        self.b_pack: Optional[torch.Tensor] = None  # [out, 64] cols >= r are zero
        self._packed_versions = (-1, -1)

    # -- reference API ---------------------------------------------------------------------------
    def apply_to(self):
        self.org_forward = self.org_module[0].forward
        self.org_module[0].forward = self.forward
        self.org_module[0]._b200_lora = weakref.ref(self)

    def _set_runtime_scale(self, value):
        self.scale = float(value)
        with torch.no_grad():
            self._runtime_scale.fill_(self.scale)

    def enable_gradient_checkpointing(self):
        self.is_checkpointing = True

    def disable_gradient_checkpointing(self):
        self.is_checkpointing = False

    def down_weight_2d(self):
        return self.lora_down.weight.view(self.lora_dim, self.in_dim)

    def up_weight_2d(self):
        return self.lora_up.weight.view(self.out_dim, self.lora_dim)

    def is_live(self) -> bool:
        """False when the reference's forward would fall through to `org_forward` (network_mixins.py:281-295)."""
        net = self.network_ref()
        if net is None or not net.is_active or net.is_merged_in:
            return False
        m = net._multiplier
        if isinstance(m, (int, float)) and m == 0:
            return False
        return True

    def has_dropout(self) -> bool:
        """dropout / rank_dropout draw masks only in training mode (network_mixins.py:211-226)."""
        return self.training and bool((self.dropout and self.dropout > 0) or (self.rank_dropout and self.rank_dropout > 0))

    def forward(self, x, *args, **kwargs):
        if not self.is_live():
            return self.org_forward(x, *args, **kwargs)
        # module dropout (network_mixins.py:198-201): the adapter's contribution is 0.0 for this call
        if self.module_dropout is not None and self.training and torch.rand(1) < self.module_dropout:
            return self.org_forward(x, *args, **kwargs)
        from .autograd import lora_linear  # late import: needs the CUDA library only when actually active

        return lora_linear(self, x)

    @torch.no_grad()
    def merge_out(self, merge_out_weight=1.0):
        self.merge_in(merge_weight=-abs(merge_out_weight))

    @torch.no_grad()
    def merge_in(self, merge_weight=1.0):
        """W += merge_weight * scale * (up @ down)  (network_mixins.py:370-462, linear / 1x1-conv cases)."""
        if not self.can_merge_in:
            return
        up = self.up_weight_2d().clone().float()
        down = self.down_weight_2d().clone().float()
        if not up.any() or not down.any():
            return
        om = self.org_module[0]
        w = om.weight
        delta = (up @ down) * (merge_weight * self.scale)
        om.weight.data = (w.float().view(self.out_dim, self.in_dim) + delta.to(w.device)).view_as(w).to(w.dtype)

    @torch.no_grad()
    def reset_weights(self):
        self.lora_up.weight.zero_()

    @torch.no_grad()
    def extract_weight(self, extract_mode="existing", extract_mode_param=None, _reflatten=True):
        """LoRM-style extraction (network_mixins.py:113-168 + toolkit/lorm.py:210-260): truncated SVD of the wrapped layer's
        OWN weight -> lora_up = U_r diag(S_r), lora_down = Vh_r; the adapter takes the extracted rank, alpha = rank, and
        the runtime scale is re-synchronised (pinned by testing/test_lora_compile_scalars.py:72-92).  Set-up-time host
        mathematics (torch.linalg.svd), not hot path."""
        if extract_mode == "existing":
            extract_mode, extract_mode_param = "fixed", self.lora_dim
        w = self.org_module[0].weight.detach().clone().float()
        out_ch = w.shape[0]
        w2 = w.reshape(out_ch, -1)
        in_ch = w2.shape[1]
        U, S, Vh = torch.linalg.svd(w2)
        if extract_mode == "percentage":
            rank = int(float(extract_mode_param) * out_ch * in_ch / (in_ch + out_ch))
        elif extract_mode == "fixed":
            rank = int(extract_mode_param)
        elif extract_mode == "threshold":
            rank = int((S > float(extract_mode_param)).sum())
        elif extract_mode == "ratio":
            rank = int((S > S.max() * float(extract_mode_param)).sum())
        elif extract_mode == "quantile":
            rank = int((torch.cumsum(S, 0) < float(extract_mode_param) * S.sum()).sum())
        else:
            raise NotImplementedError('Extract mode should be "fixed", "threshold", "ratio" or "quantile"')
        rank = min(out_ch, in_ch, max(1, rank))
        if rank >= out_ch / 2:
            rank = int(out_ch / 2)
        if rank > RANK_PAD:
            raise NotImplementedError(f"extracted rank {rank} > {RANK_PAD}")
        dev = self.lora_down.weight.device
        up = (U[:, :rank] @ torch.diag(S[:rank])).reshape(out_ch, rank)
        down = Vh[:rank, :].reshape(rank, in_ch)
        self.lora_dim = rank
        if self.is_conv:
            down = down.view(rank, self.in_channels, *self.kernel_size)
            up = up.view(out_ch, rank, 1, 1)
        self.lora_down.weight = nn.Parameter(down.to(dev, torch.float32).clone())
        self.lora_up.weight = nn.Parameter(up.to(dev, torch.float32).clone())
        self.alpha = (self.alpha * 0) + rank
        self._set_runtime_scale(float(self.alpha.detach().float().item()) / self.lora_dim)
        net = self.network_ref() if self.network_ref is not None else None
        if _reflatten and net is not None and net.flat_params is not None:
            net._flatten()  # shapes changed: rebuild the flat views and operand packs


class FusedGroup:
    """Adapters that share one input (to_q / to_k / to_v, ...): fused bf16 operands for one GEMM over the group."""

    def __init__(self, loras, a_fused, b_fused):
        self.loras, self.a_fused, self.b_fused = loras, a_fused, b_fused
        self.r = loras[0].lora_dim
        self.out_dims = [m.out_dim for m in loras]


class LoRASpecialNetwork(nn.Module):
    NUM_OF_BLOCKS = 12
    UNET_TARGET_REPLACE_MODULE = ["UNet2DConditionModel"]
    UNET_TARGET_REPLACE_MODULE_CONV2D_3X3 = ["UNet2DConditionModel"]
    TEXT_ENCODER_TARGET_REPLACE_MODULE = ["CLIPAttention", "CLIPMLP"]
    LORA_PREFIX_UNET = "lora_unet"
    PEFT_PREFIX_UNET = "unet"
    LORA_PREFIX_TEXT_ENCODER = "lora_te"
    LORA_PREFIX_TEXT_ENCODER1 = "lora_te1"
    LORA_PREFIX_TEXT_ENCODER2 = "lora_te2"

    def __init__(self, text_encoder, unet, multiplier: float = 1.0, lora_dim: int = 4, alpha: float = 1,
                 dropout: Optional[float] = None, rank_dropout: Optional[float] = None,
                 module_dropout: Optional[float] = None, conv_lora_dim: Optional[int] = None,
                 conv_alpha: Optional[float] = None, block_dims=None, block_alphas=None, conv_block_dims=None,
                 conv_block_alphas=None, modules_dim: Optional[Dict[str, int]] = None,
                 modules_alpha: Optional[Dict[str, int]] = None, module_class=LoRAModule, varbose: Optional[bool] = False,
                 train_text_encoder: Optional[bool] = True, use_text_encoder_1: bool = True, use_text_encoder_2: bool = True,
                 train_unet: Optional[bool] = True, is_sdxl=False, is_v2=False, is_v3=False, is_pixart: bool = False,
                 is_auraflow: bool = False, is_flux: bool = False, is_lumina2: bool = False, use_bias: bool = False,
                 is_lorm: bool = False, ignore_if_contains=None, only_if_contains=None, full_if_contains=None,
                 parameter_threshold: float = 0.0, attn_only: bool = False, target_lin_modules=None,
                 target_conv_modules=None, network_type: str = "lora", full_train_in_out: bool = False,
                 transformer_only: bool = False, peft_format: bool = False, is_assistant_adapter: bool = False,
                 is_transformer: bool = False, base_model=None, is_ara: bool = False, is_ssd=False, is_vega=False,
                 network_config=None, **kwargs) -> None:
        super().__init__()
        if network_type.lower() != "lora":
            raise NotImplementedError(f"network_type={network_type!r}: only 'lora' is implemented on the B200 path "
                                      "(DoRA / LoKr / fullrank are SURVEY.md section 8(f) rows)")
        if is_lorm or full_train_in_out or (full_if_contains is not None and len(full_if_contains) > 0):
            raise NotImplementedError("LoRM / full_train_in_out / full_if_contains are not implemented")
        # --- ToolkitNetworkMixin.__init__ (network_mixins.py:491-523)
        self.train_text_encoder = train_text_encoder
        self.train_unet = train_unet
        self.is_checkpointing = False
        self._multiplier = 1.0
        self.is_active = False
        self.is_sdxl, self.is_ssd, self.is_vega, self.is_v2 = is_sdxl, is_ssd, is_vega, is_v2
        self.is_v1 = not is_v2 and not is_sdxl and not is_ssd and not is_vega
        self.is_merged_in = False
        self.is_lorm = is_lorm
        self.network_config = network_config
        self.module_losses: List[torch.Tensor] = []
        self.can_merge_in = True
        self.did_change_weights = False
        # --- LoRASpecialNetwork.__init__ (lora_special.py:343-433)
        self.ignore_if_contains = ignore_if_contains if ignore_if_contains is not None else []
        self.only_if_contains = only_if_contains
        self.full_if_contains = []
        self.base_model_ref = weakref.ref(base_model) if base_model is not None else None
        self.lora_dim = lora_dim
        self.alpha = alpha
        self.conv_lora_dim = conv_lora_dim
        self.conv_alpha = conv_alpha
        self.dropout, self.rank_dropout, self.module_dropout = dropout, rank_dropout, module_dropout
        self.is_v3, self.is_pixart, self.is_auraflow = is_v3, is_pixart, is_auraflow
        self.is_flux, self.is_lumina2 = is_flux, is_lumina2
        self.network_type = network_type
        self.is_assistant_adapter = is_assistant_adapter
        self.full_rank = False
        self.is_ara = is_ara
        self.transformer_only = transformer_only
        self.peft_format = peft_format
        self.is_transformer = is_transformer
        self.full_train_in_out = False
        self.use_old_lokr_format = False
        if self.is_flux or self.is_v3 or self.is_lumina2 or is_transformer:
            self.peft_format = True
        if self.peft_format:  # "no alpha for peft" (lora_special.py:428-433)
            self.alpha = self.lora_dim
            alpha = self.alpha
            self.conv_alpha = self.conv_lora_dim
            conv_alpha = self.conv_alpha
        # the reference's defaults are kohya's class attributes, not this class's (lora_special.py:331-332,
        # kohya_lora.py:750-751): LoRA goes on the Linear / 1x1-conv layers inside `Transformer2DModel` blocks
        if target_lin_modules is None:
            target_lin_modules = list(KOHYA_UNET_TARGET_REPLACE_MODULE)

        def create_modules(is_unet, text_encoder_idx, root_module, target_replace_modules):
            unet_prefix = self.PEFT_PREFIX_UNET if self.peft_format else self.LORA_PREFIX_UNET
            if is_pixart or is_v3 or is_auraflow or is_flux or is_lumina2 or self.is_transformer:
                unet_prefix = "transformer" if self.peft_format else "lora_transformer"
            prefix = unet_prefix if is_unet else (
                self.LORA_PREFIX_TEXT_ENCODER if text_encoder_idx is None else
                (self.LORA_PREFIX_TEXT_ENCODER1 if text_encoder_idx == 1 else self.LORA_PREFIX_TEXT_ENCODER2))
            loras, skipped = [], []
            for name, module in root_module.named_modules():
                if module.__class__.__name__ not in target_replace_modules:
                    continue
                for child_name, child_module in module.named_modules():
                    cls = child_module.__class__.__name__
                    is_linear = cls in LINEAR_MODULES
                    is_conv2d = cls in CONV_MODULES
                    is_conv2d_1x1 = is_conv2d and tuple(child_module.kernel_size) == (1, 1)
                    parts = [x for x in (prefix, name, child_name) if x]
                    clean_name = ".".join(parts)
                    lora_name = clean_name.replace(".", "$$") if self.peft_format else clean_name.replace(".", "_")
                    skip = any(word in clean_name for word in self.ignore_if_contains)
                    if sum(p.numel() for p in child_module.parameters()) < parameter_threshold:
                        skip = True
                    if self.transformer_only and is_unet:
                        block_names = base_model.get_transformer_block_names() if base_model is not None else None
                        if block_names is not None:
                            if not any(b in clean_name for b in block_names):
                                skip = True
                        else:
                            if (self.is_pixart or self.is_flux or self.is_v3) and "transformer_blocks" not in lora_name:
                                skip = True
                            if self.is_lumina2 and not any(s in lora_name for s in ("layers$$", "noise_refiner$$",
                                                                                      "context_refiner$$")):
                                skip = True
                            if hasattr(root_module, "transformer_blocks") and "transformer_blocks" not in lora_name:
                                skip = True
                            if hasattr(root_module, "blocks") and "blocks" not in lora_name:
                                skip = True
                            if hasattr(root_module, "single_blocks") and "single_blocks" not in lora_name \
                                    and "double_blocks" not in lora_name:
                                skip = True
                    if not (is_linear or is_conv2d) or skip:
                        continue
                    if self.only_if_contains is not None:
                        if not any(w in clean_name for w in self.only_if_contains) and \
                                not any(w in lora_name for w in self.only_if_contains):
                            continue
                    dim = alpha_ = None
                    if modules_dim is not None:
                        if lora_name in modules_dim:
                            dim, alpha_ = modules_dim[lora_name], modules_alpha[lora_name]
                    elif is_linear or is_conv2d_1x1:
                        dim, alpha_ = self.lora_dim, self.alpha
                    elif self.conv_lora_dim is not None:  # k x k convs (lora_special.py:585-587)
                        dim, alpha_ = self.conv_lora_dim, self.conv_alpha
                    if dim is None or dim == 0:
                        if is_linear or is_conv2d_1x1 or self.conv_lora_dim is not None or conv_block_dims is not None:
                            skipped.append(lora_name)
                        continue
                    loras.append(module_class(lora_name, child_module, self.multiplier, dim, alpha_, dropout=dropout,
                                              rank_dropout=rank_dropout, module_dropout=module_dropout, network=self,
                                              parent=module, use_bias=use_bias))
            return loras, skipped

        text_encoders = text_encoder if isinstance(text_encoder, list) else [text_encoder]
        self.text_encoder_loras = []
        if train_text_encoder:
            for i, te in enumerate(text_encoders):
                if te is None or (not use_text_encoder_1 and i == 0) or (not use_text_encoder_2 and i == 1):
                    continue
                index = i + 1 if len(text_encoders) > 1 else None
                replace = ["T5EncoderModel"] if self.is_pixart else self.TEXT_ENCODER_TARGET_REPLACE_MODULE
                te_loras, _ = create_modules(False, index, te, replace)
                self.text_encoder_loras.extend(te_loras)
        target_modules = list(target_lin_modules)
        if modules_dim is not None or self.conv_lora_dim is not None or conv_block_dims is not None:  # :679-681
            target_modules += list(target_conv_modules or KOHYA_UNET_TARGET_REPLACE_MODULE_CONV2D_3X3)
        if is_v3:
            target_modules = ["SD3Transformer2DModel"]
        if is_pixart:
            target_modules = ["PixArtTransformer2DModel"]
        if is_auraflow:
            target_modules = ["AuraFlowTransformer2DModel"]
        if is_flux:
            target_modules = ["FluxTransformer2DModel"]
        if is_lumina2:
            target_modules = ["Lumina2Transformer2DModel"]
        self.unet_loras = create_modules(True, None, unet, target_modules)[0] if train_unet and unet is not None else []
        print(f"create LoRA network (b200). base dim (rank): {lora_dim}, alpha: {alpha}; "
              f"text encoder: {len(self.text_encoder_loras)} modules, U-Net/transformer: {len(self.unet_loras)} modules.")
        self.up_lr_weight = self.down_lr_weight = self.mid_lr_weight = None
        self.block_lr = False
        names = set()
        for lora in self.text_encoder_loras + self.unet_loras:
            assert lora.lora_name not in names, f"duplicated lora name: {lora.lora_name}"
            names.add(lora.lora_name)
        # flat storage (filled by _flatten once the modules are registered / moved)
        self.flat_params: Optional[torch.Tensor] = None
        self.flat_grads: Optional[torch.Tensor] = None
        self.pack_buf: Optional[torch.Tensor] = None
        self._repack_table = None
        self._fused_group_defs = []
        self.fused_groups = {}
        self.torch_multiplier = torch.tensor((1.0,))
        self.multiplier = multiplier

    # ------------------------------------------------------------------------------------------
    # flat parameter storage
    # ------------------------------------------------------------------------------------------
    def get_all_modules(self) -> List[LoRAModule]:
        return list(self.unet_loras) + list(self.text_encoder_loras)

    @torch.no_grad()
    def _flatten(self):
        """(Re)build the flat fp32 parameter / gradient buffers and the bf16 operand pack on the modules' device."""
        mods = self.get_all_modules()
        if not mods:
            return
        dev = mods[0].lora_down.weight.device
        total = sum(m.lora_down.weight.numel() + m.lora_up.weight.numel() for m in mods)
        total_pad = (total + 3) // 4 * 4
        flat = torch.zeros(total_pad, device=dev, dtype=torch.float32)
        grads = torch.zeros(total_pad, device=dev, dtype=torch.float32)
        off = 0
        self._param_offsets = {}
        for m in mods:
            for lin in (m.lora_down, m.lora_up):
                n = lin.weight.numel()
                flat[off:off + n].copy_(lin.weight.detach().reshape(-1).to(dev, torch.float32))
                lin.weight.data = flat[off:off + n].view(lin.weight.shape)
                lin.weight.grad = grads[off:off + n].view(lin.weight.shape)
                self._param_offsets[id(lin)] = off
                off += n
        self.flat_params, self.flat_grads = flat, grads
        self.n_params = total
        self._build_packs()

    @torch.no_grad()
    def _build_packs(self):
        """bf16 operand buffers: per module A_pack [64, in] / B_pack [out, 64], plus one (A_fused [64, in],
        B_fused [sum out, 64]) pair per fused group (adapters sharing an input, e.g. to_q/to_k/to_v: module j's A goes
        to rows [j r, (j+1) r) of A_fused and its B to the block (rows of module j, columns [j r, (j+1) r)) of B_fused,
        so ONE rank-side GEMM and ONE fused GEMM serve the whole group).  One repack table covers everything."""
        mods = self.get_all_modules()
        dev = self.flat_params.device
        groups = [g for g in self._fused_group_defs]
        pack_elems = sum(RANK_PAD * m.in_dim + m.out_dim * RANK_PAD for m in mods)
        pack_elems += sum(RANK_PAD * g[0].in_dim + sum(m.out_dim for m in g) * RANK_PAD for g in groups)
        pack = torch.zeros(pack_elems, device=dev, dtype=torch.bfloat16)
        entries = []
        poff = 0
        for m in mods:
            oa, ob = self._param_offsets[id(m.lora_down)], self._param_offsets[id(m.lora_up)]
            m.a_pack = pack[poff:poff + RANK_PAD * m.in_dim].view(RANK_PAD, m.in_dim)
            entries.append(cabi.RepackEntry(oa, poff, m.lora_dim, m.in_dim, m.in_dim, 0))
            poff += RANK_PAD * m.in_dim
            m.b_pack = pack[poff:poff + m.out_dim * RANK_PAD].view(m.out_dim, RANK_PAD)
            entries.append(cabi.RepackEntry(ob, poff, m.out_dim, m.lora_dim, RANK_PAD, 0))
            poff += m.out_dim * RANK_PAD
            m._packed_versions = (-1, -1)
        self.fused_groups = {}
        for g in groups:
            r, k = g[0].lora_dim, g[0].in_dim
            n_out = sum(m.out_dim for m in g)
            a_f = pack[poff:poff + RANK_PAD * k].view(RANK_PAD, k)
            a_off = poff
            poff += RANK_PAD * k
            b_f = pack[poff:poff + n_out * RANK_PAD].view(n_out, RANK_PAD)
            b_off = poff
            poff += n_out * RANK_PAD
            row = 0
            for j, m in enumerate(g):
                entries.append(cabi.RepackEntry(self._param_offsets[id(m.lora_down)], a_off + j * r * k, r, k, k, 0))
                entries.append(cabi.RepackEntry(self._param_offsets[id(m.lora_up)], b_off + row * RANK_PAD + j * r, m.out_dim, r,
                                                RANK_PAD, 0))
                row += m.out_dim
            self.fused_groups[tuple(id(m) for m in g)] = FusedGroup(list(g), a_f, b_f)
        table = (cabi.RepackEntry * len(entries))(*entries)
        self.pack_buf = pack
        self._repack_entries = len(entries)
        self._repack_table_host = table
        self._repack_table = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(dev) if dev.type == "cuda" else None
        self._pack_dirty = True

    def register_fused_groups(self, groups):
        """Register lists of adapters that share one input; invalid groups (different input / rank / scale, rank not a
        multiple of 8, more than 64 rank columns in total, a missing adapter) are skipped.  One pack rebuild."""
        added = False
        for loras in groups:
            if any(l is None for l in loras):
                continue
            key = tuple(id(m) for m in loras)
            if key in self.fused_groups or any(tuple(id(m) for m in g) == key for g in self._fused_group_defs):
                continue
            r, k = loras[0].lora_dim, loras[0].in_dim
            if any(m.lora_dim != r or m.in_dim != k or m.scale != loras[0].scale for m in loras) or r % 8 != 0 \
                    or r * len(loras) > RANK_PAD:
                continue
            self._fused_group_defs.append(list(loras))
            added = True
        if added:
            self._build_packs()

    def fused_group(self, loras):
        """The registered FusedGroup of these adapters, or None."""
        if any(l is None for l in loras):
            return None
        return self.fused_groups.get(tuple(id(m) for m in loras))

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        # .to() / .cuda() / .float() replaced every parameter tensor: rebuild the flat views on the new device.  Before
        # apply_to() the adapters are not registered sub-modules yet (force_to moves them itself), so nothing to do.
        mods = self.get_all_modules()
        if mods and all(m.lora_name in self._modules for m in mods):
            for m in mods:
                if m.lora_down.weight.dtype != torch.float32:
                    raise NotImplementedError("LoRA master weights must be fp32 (BaseSDTrainProcess.py:1983)")
            self._flatten()
        return out

    def ensure_grad_views(self):
        """Re-attach `.grad` views (and zero them) after an external `zero_grad(set_to_none=True)`."""
        mods = self.get_all_modules()
        if not mods or self.flat_grads is None:
            return
        w = mods[0].lora_down.weight
        if w.grad is not None and w.grad.data_ptr() == self.flat_grads.data_ptr():
            return
        self.flat_grads.zero_()
        off = 0
        for m in mods:
            for lin in (m.lora_down, m.lora_up):
                n = lin.weight.numel()
                lin.weight.grad = self.flat_grads[off:off + n].view(lin.weight.shape)
                off += n

    def mark_params_changed(self):
        self._pack_dirty = True

    def refresh_packs(self, force=False):
        """bf16 operand copies of A / B for the tensor cores (one launch over all modules)."""
        if self.flat_params is None:
            self._flatten()
        mods = self.get_all_modules()
        if not force and not self._pack_dirty:
            # parameters may have been updated by a foreign optimizer: compare tensor versions of one module
            m = mods[0]
            if m._packed_versions == (m.lora_down.weight._version, m.lora_up.weight._version):
                return
        if self.flat_params.device.type != "cuda":
            raise cabi.B200Error("LoRASpecialNetwork is active on a non-CUDA device; the B200 path has no CPU fallback")
        from . import ops

        ops.repack_lora(self.flat_params, self.pack_buf, self._repack_table, self._repack_entries)
        for m in mods:
            m._packed_versions = (m.lora_down.weight._version, m.lora_up.weight._version)
        self._pack_dirty = False

    # ------------------------------------------------------------------------------------------
    # reference API (kohya_lora.py / network_mixins.py)
    # ------------------------------------------------------------------------------------------
    def apply_to(self, text_encoder, unet, apply_text_encoder=True, apply_unet=True):
        if not apply_text_encoder:
            self.text_encoder_loras = []
        if not apply_unet:
            self.unet_loras = []
        for lora in self.text_encoder_loras + self.unet_loras:
            lora.apply_to()
            self.add_module(lora.lora_name, lora)
        # the fused engines find their network through the model, not through one particular Linear (module-selection
        # options such as only_if_contains / ignore_if_contains may leave any given Linear unadapted)
        for root in ([unet] if unet is not None and self.unet_loras else []):
            try:
                root._b200_network = weakref.ref(self)
            except Exception:  # objects that refuse attributes: the engines then scan for an adapted Linear
                pass
        self._flatten()

    def is_mergeable(self):
        return True

    def prepare_optimizer_params(self, text_encoder_lr, unet_lr, default_lr):
        self.requires_grad_(True)
        all_params = []

        def enumerate_params(loras):
            params = []
            for lora in loras:
                params.extend(lora.parameters())
            return params

        if self.text_encoder_loras:
            pd = {"params": enumerate_params(self.text_encoder_loras)}
            if text_encoder_lr is not None:
                pd["lr"] = text_encoder_lr
            all_params.append(pd)
        if self.unet_loras:
            pd = {"params": enumerate_params(self.unet_loras)}
            if unet_lr is not None:
                pd["lr"] = unet_lr
            all_params.append(pd)
        return all_params

    def prepare_grad_etc(self, text_encoder=None, unet=None):
        self.requires_grad_(True)

    def on_epoch_start(self, text_encoder=None, unet=None):
        self.train()

    def get_trainable_params(self):
        return self.parameters()

    def enable_gradient_checkpointing(self):
        self.is_checkpointing = True
        for m in self.get_all_modules():
            m.enable_gradient_checkpointing()

    def disable_gradient_checkpointing(self):
        self.is_checkpointing = False
        for m in self.get_all_modules():
            m.disable_gradient_checkpointing()

    @torch.no_grad()
    def _update_torch_multiplier(self):
        mods = self.get_all_modules()
        if not mods:
            raise ValueError("There are not any lora modules in this network. Check your config and try again")
        w = mods[0].lora_down.weight
        m = self._multiplier
        if isinstance(m, (int, float)):
            t = torch.tensor((m,))
        elif isinstance(m, list):
            t = torch.tensor(m)
        else:
            t = m.clone().detach()
        self.torch_multiplier = t.to(w.device, dtype=w.dtype).clone().detach()

    @property
    def multiplier(self):
        return self._multiplier

    @multiplier.setter
    def multiplier(self, value):
        same = False
        try:
            same = bool(self._multiplier == value)
        except Exception:
            same = False
        if same:
            return
        self._multiplier = value
        if self.get_all_modules():
            self._update_torch_multiplier()

    def __enter__(self):
        self.is_active = True

    def __exit__(self, exc_type, exc_value, tb):
        self.is_active = False

    def force_to(self, device, dtype):
        self.to(device, dtype)
        for lora in self.get_all_modules():  # not registered as sub-modules before apply_to (network_mixins.py:855-863)
            lora.to(device, dtype)

    def reset_weights(self):
        for m in self.get_all_modules():
            m.reset_weights()
        self.mark_params_changed()

    def merge_in(self, merge_weight=1.0):
        self.is_merged_in = True
        for m in self.get_all_modules():
            m.merge_in(merge_weight)

    def merge_out(self, merge_weight=1.0):
        if not self.is_merged_in:
            return
        self.is_merged_in = False
        for m in self.get_all_modules():
            m.merge_out(merge_weight)

    def extract_weight(self, extract_mode="existing", extract_mode_param=None):
        """network_mixins.py:908-919: every adapter re-initialised from the SVD of its wrapped layer."""
        if extract_mode_param is None:
            raise ValueError("extract_mode_param must be set")
        for m in self.get_all_modules():
            m.extract_weight(extract_mode=extract_mode, extract_mode_param=extract_mode_param, _reflatten=False)
        if self.flat_params is not None:
            self._flatten()

    # -- state dict / files (network_mixins.py:525-789) ------------------------------------------
    def get_keymap(self, force_weight_mapping=False):
        from .keymaps import load_keymap

        return load_keymap(self, force_weight_mapping)

    def get_state_dict(self, extra_state_dict=None, dtype=torch.float16):
        keymap = self.get_keymap()
        save_keymap = {v: k for k, v in keymap.items()} if keymap is not None else {}
        save_dict = OrderedDict()
        for key, v in self.state_dict().items():
            save_dict[save_keymap.get(key, key)] = v.detach().clone().to("cpu").to(dtype)
        if extra_state_dict is not None:
            for key, v in extra_state_dict.items():
                save_dict[key] = v.detach().clone().to("cpu").to(dtype)
        if self.peft_format:
            new = {}
            for key, value in save_dict.items():
                if key.endswith(".alpha"):
                    continue
                new[key.replace("lora_down", "lora_A").replace("lora_up", "lora_B").replace("$$", ".")] = value
            save_dict = new
        if self.base_model_ref is not None and self.base_model_ref() is not None:
            save_dict = self.base_model_ref().convert_lora_weights_before_save(save_dict)
        return save_dict

    def save_weights(self, file, dtype=torch.float16, metadata=None, extra_state_dict=None):
        from .metadata import add_model_hash_to_meta

        save_dict = self.get_state_dict(extra_state_dict=extra_state_dict, dtype=dtype)
        if metadata is not None and len(metadata) == 0:
            metadata = None
        if metadata is None:
            metadata = OrderedDict()
        metadata = add_model_hash_to_meta(save_dict, metadata)
        base = self.base_model_ref() if self.base_model_ref is not None else None
        if base is not None and hasattr(base, "save_lora"):
            base.save_lora(save_dict, file, metadata)
            return
        if os.path.splitext(file)[1] == ".safetensors":
            from safetensors.torch import save_file

            save_file({k: v.contiguous() for k, v in save_dict.items()}, file, metadata)
        else:
            torch.save(save_dict, file)

    def load_weights(self, file, force_weight_mapping=False):
        keymap = self.get_keymap(force_weight_mapping) or {}
        base = self.base_model_ref() if self.base_model_ref is not None else None
        if isinstance(file, str):
            if base is not None and hasattr(base, "load_lora"):
                weights_sd = base.load_lora(file)
            elif os.path.splitext(file)[1] == ".safetensors":
                from safetensors.torch import load_file

                weights_sd = load_file(file)
            else:
                weights_sd = torch.load(file, map_location="cpu")
        else:
            weights_sd = file
        if base is not None:
            weights_sd = base.convert_lora_weights_before_load(weights_sd)
        load_sd = OrderedDict()
        for key, value in weights_sd.items():
            load_key = keymap.get(key, key)
            if self.is_pixart:
                load_key = load_key.replace("__", "_")
            if self.peft_format:
                if load_key.endswith(".alpha"):
                    continue
                load_key = load_key.replace("lora_A", "lora_down").replace("lora_B", "lora_up")
                load_key = load_key.replace(".", "$$")
                load_key = load_key.replace("$$lora_down$$", ".lora_down.").replace("$$lora_up$$", ".lora_up.")
            load_sd[load_key] = value
        current = self.state_dict()
        extra_dict, to_delete = OrderedDict(), []
        for key in list(load_sd.keys()):
            if key not in current:
                extra_dict[key] = load_sd[key]
                to_delete.append(key)
            elif ("lora_down" in key or "lora_up" in key) and load_sd[key].dim() == 2:
                lv, tgt = load_sd[key], current[key]
                (th, tw), (sh, sw) = tgt.shape, lv.shape
                if (sh, sw) == (th, tw):
                    pass
                elif ("lora_down" in key and sh < th) or ("lora_up" in key and sw < tw):  # expand rank
                    nv = torch.zeros((th, tw), device=lv.device, dtype=lv.dtype)
                    nv[:sh, :sw] = lv
                    load_sd[key] = nv
                    self.did_change_weights = True
                elif ("lora_down" in key and sh > th) or ("lora_up" in key and sw > tw):  # shrink rank
                    load_sd[key] = lv[:th, :tw]
                    self.did_change_weights = True
                else:
                    raise ValueError(f"Unhandled LoRA shape change for {key}: src={lv.shape}, tgt={tgt.shape}")
        for key in to_delete:
            del load_sd[key]
        print(f"Missing keys: {to_delete}")
        if len(to_delete) > 0 and self.is_v1 and not force_weight_mapping and not (
                len(to_delete) == 1 and "emb_params" in to_delete):
            print(" Attempting to load with forced keymap")
            return self.load_weights(file, force_weight_mapping=True)
        # copy INTO the flat views (load_state_dict copies in place, so the views stay bound)
        self.load_state_dict(load_sd, False)
        self.mark_params_changed()
        return extra_dict if len(extra_dict) > 0 else None


def get_network(unet, text_encoder=None, *, network_config=None, model_config=None, train_config=None, base_model=None,
                device=None, **network_kwargs) -> LoRASpecialNetwork:
    """The inline factory of `jobs/process/BaseSDTrainProcess.py:1926-1993` as a function.

    `network_config` / `model_config` / `train_config` are duck-typed (attributes as in
    `toolkit/config_modules.py:169-202, 377-417`); keyword overrides win.  Returns the network already
    `force_to(device, fp32)`, `_update_torch_multiplier()`-ed and `apply_to`-ed, as the trainer does.
    """
    def g(obj, name, default):
        return getattr(obj, name, default) if obj is not None else default

    kw = dict(
        text_encoder=text_encoder, unet=unet, lora_dim=g(network_config, "linear", 4), multiplier=1.0,
        alpha=g(network_config, "linear_alpha", 1.0), train_unet=g(train_config, "train_unet", True),
        train_text_encoder=g(train_config, "train_text_encoder", False), conv_lora_dim=g(network_config, "conv", None),
        conv_alpha=g(network_config, "conv_alpha", None), is_sdxl=g(model_config, "is_xl", False) or g(model_config, "is_ssd", False),
        is_v2=g(model_config, "is_v2", False), is_v3=g(model_config, "is_v3", False),
        is_pixart=g(model_config, "is_pixart", False), is_auraflow=g(model_config, "is_auraflow", False),
        is_flux=g(model_config, "is_flux", False), is_lumina2=g(model_config, "is_lumina2", False),
        is_ssd=g(model_config, "is_ssd", False), is_vega=g(model_config, "is_vega", False),
        dropout=g(network_config, "dropout", None), use_text_encoder_1=g(model_config, "use_text_encoder_1", True),
        use_text_encoder_2=g(model_config, "use_text_encoder_2", True), use_bias=False, is_lorm=False,
        network_config=network_config, network_type=g(network_config, "type", "lora"),
        transformer_only=g(network_config, "transformer_only", True), is_transformer=g(base_model, "is_transformer", False),
        base_model=base_model)
    kw.update(g(network_config, "network_kwargs", {}) or {})
    if base_model is not None and hasattr(base_model, "target_lora_modules"):  # :1945-1946
        kw["target_lin_modules"] = base_model.target_lora_modules
    kw.update(network_kwargs)
    net = LoRASpecialNetwork(**kw)
    if device is None:
        device = next(unet.parameters()).device
    net.force_to(device, dtype=torch.float32)
    if base_model is not None:  # "give network to sd so it can use it" (:1984)
        try:
            base_model.network = net
        except AttributeError:
            pass
    net._update_torch_multiplier()
    net.apply_to(text_encoder, unet, kw["train_text_encoder"], kw["train_unet"])
    return net
