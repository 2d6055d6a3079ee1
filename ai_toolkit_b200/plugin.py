"""The plugin seam: `SDTrainerB200`, an `SDTrainer` whose `hook_train_loop` is the fused B200 step.

Reference interface: a package under `extensions/` exporting `AI_TOOLKIT_EXTENSIONS = [Extension]`
(toolkit/extension.py:9-57, extensions/example/__init__.py); the process class is constructed by the job
(`jobs/BaseJob.py:57-66`) and driven by `BaseSDTrainProcess.run()` (jobs/process/BaseSDTrainProcess.py:1745-2848), which
calls the hooks overridden here:

  hook_before_model_load (:1752)   `BaseSDTrainProcess.run` builds its network from the module-level name
                                   `LoRASpecialNetwork` (:1949): re-pointed at this package's class for network type 'lora'
  hook_after_model_load (:1923)    the diffusers FluxTransformer2DModel is replaced by this package's parameter container
                                   (same parameter names, SAME storage: no second copy of the 23.8 GB of weights)
  hook_before_train_loop (:2260)   torch.optim.AdamW -> B200AdamW (state carried over), lr scheduler re-bound to it,
                                   ExponentialMovingAverage -> FusedEMA (the EMA lives in the AdamW kernel)
  hook_train_loop (:2624-2625)     SDTrainer.hook_train_loop (extensions_built_in/sd_trainer/SDTrainer.py:2243-2317) with
                                   train_single_accumulation + clip_grad_norm_ + optimizer.step + ema.update replaced by
                                   `FluxLoRATrainStep`; batch preparation stays the reference's own
                                   `process_general_training_batch` (all its config switches keep working)

The stock `hook_train_loop` must NOT run on top of this one: it would clip and EMA-update a second time.  This module
imports nothing from the reference at import time; `make_trainer_class(SDTrainer)` builds the subclass.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

UID = "sd_trainer_b200"
NAME = "SD Trainer (B200 fused LoRA path)"


class FusedEMA:
    """What the trainer still needs from `toolkit.ema.ExponentialMovingAverage` once the update itself is fused into the
    clip/AdamW kernel: `eval()` swaps the EMA weights in for sampling / saving (BaseSDTrainProcess.py:370-385, :509-511,
    :719-720), `train()` swaps the training weights back, `update()` is a no-op kept for callers that still call it."""

    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.decay = optimizer.ema_decay
        self._is_train_mode = True
        self._stash = None

    def update(self, *a, **k):  # fused: shadow -= (1 - decay)(shadow - p) happens inside b200_clip_adamw
        return None

    @torch.no_grad()
    def eval(self):
        opt = self.optimizer
        if self._is_train_mode and opt.ema is not None:
            net = opt.network
            self._stash = net.flat_params.clone()
            net.flat_params.copy_(opt.ema)
            net.mark_params_changed()
            self._is_train_mode = False

    @torch.no_grad()
    def train(self):
        opt = self.optimizer
        if not self._is_train_mode and self._stash is not None:
            opt.network.flat_params.copy_(self._stash)
            opt.network.mark_params_changed()
            self._stash = None
        self._is_train_mode = True

    def state_dict(self):
        return {"decay": self.decay, "shadow_flat": None if self.optimizer.ema is None else self.optimizer.ema.detach().cpu()}

    def load_state_dict(self, sd):
        if sd.get("shadow_flat") is not None and self.optimizer.ema is not None:
            self.optimizer.ema.copy_(sd["shadow_flat"])


def adopt_flux_transformer(src):
    """diffusers `FluxTransformer2DModel` -> this package's container over the SAME parameter storage
    (`load_state_dict(assign=True)`): module paths / parameter names are identical, so LoRA names and checkpoints are too."""
    from .flux import FluxConfig, FluxTransformer2DModel

    if isinstance(src, FluxTransformer2DModel):
        return src
    c = getattr(src, "config", None)
    get = (lambda k, d: getattr(c, k, d)) if c is not None and not isinstance(c, dict) else (lambda k, d: (c or {}).get(k, d))
    cfg = FluxConfig(in_channels=get("in_channels", 64), num_layers=get("num_layers", 19),
                     num_single_layers=get("num_single_layers", 38), attention_head_dim=get("attention_head_dim", 128),
                     num_attention_heads=get("num_attention_heads", 24), joint_attention_dim=get("joint_attention_dim", 4096),
                     pooled_projection_dim=get("pooled_projection_dim", 768), guidance_embeds=get("guidance_embeds", True),
                     axes_dims_rope=tuple(get("axes_dims_rope", (16, 56, 56))))
    sd = src.state_dict()
    dtype = next(iter(sd.values())).dtype
    with torch.device("meta"):
        dst = FluxTransformer2DModel(cfg, device="meta", dtype=dtype)
    dst.load_state_dict(sd, strict=True, assign=True)
    dst.requires_grad_(False)
    return dst


def _is_unet(model) -> bool:
    return model is not None and type(model).__name__ == "UNet2DConditionModel"


def make_trainer_class(SDTrainerBase, get_lr_scheduler=None):
    """-> `SDTrainerB200(SDTrainerBase)`.  `SDTrainerBase` is the reference's
    `extensions_built_in.sd_trainer.SDTrainer.SDTrainer`; `get_lr_scheduler` its `toolkit.scheduler.get_lr_scheduler`."""

    class SDTrainerB200(SDTrainerBase):
        # ---------------------------------------------------------------- construction-time hooks
        def hook_before_model_load(self):
            super().hook_before_model_load()
            nc = getattr(self, "network_config", None)
            if nc is not None and str(getattr(nc, "type", "lora")).lower() == "lora":
                try:
                    import jobs.process.BaseSDTrainProcess as bsp  # the module whose global name run() instantiates (:1949)

                    from . import LoRASpecialNetwork

                    bsp.LoRASpecialNetwork = LoRASpecialNetwork
                except ImportError:  # not running inside ai-toolkit (tests drive the hooks directly)
                    pass

        def hook_after_model_load(self):
            super().hook_after_model_load()
            sd = self.sd
            if getattr(sd, "is_flux", False) and getattr(sd, "unet", None) is not None:
                sd.unet = adopt_flux_transformer(sd.unet)
                if getattr(sd, "pipeline", None) is not None and hasattr(sd.pipeline, "transformer"):
                    sd.pipeline.transformer = sd.unet
            elif _is_unet(getattr(sd, "unet", None)):
                # SD1.5 / SDXL: the adapter-bearing Transformer2DModel stacks move onto the engine (same tensors, same names);
                # the frozen ResNet / sampler body stays diffusers' eager model
                from .unet_blocks import adopt_unet_transformers

                adopt_unet_transformers(sd.unet)

        def hook_before_train_loop(self):
            super().hook_before_train_loop()
            self.b200_setup()

        # ---------------------------------------------------------------- optimizer / EMA / scheduler hand-over
        def b200_setup(self):
            from .optimizer import B200AdamW

            tc = self.train_config
            if self.network is None or not hasattr(self.network, "flat_params"):
                raise RuntimeError("sd_trainer_b200 needs network.type == 'lora' built by ai_toolkit_b200.LoRASpecialNetwork")
            if str(getattr(tc, "optimizer", "adamw")).lower() not in ("adamw", "adam", "adamw8bit", "adam8bit", "adamw8"):
                raise NotImplementedError(f"optimizer {tc.optimizer!r}: the fused step implements AdamW (toolkit/optimizer.py:78-79)")
            op = dict(getattr(tc, "optimizer_params", None) or {})
            old = self.optimizer
            ema_cfg = getattr(tc, "ema_config", None)
            use_ema = bool(ema_cfg is not None and getattr(ema_cfg, "use_ema", False))
            lr = old.param_groups[0]["lr"] if old is not None and len(old.param_groups) else tc.lr
            opt = B200AdamW(self.network, lr=lr, betas=tuple(op.get("betas", (0.9, 0.999))), eps=op.get("eps", 1e-6),
                            weight_decay=op.get("weight_decay", 1e-2), max_grad_norm=getattr(tc, "max_grad_norm", 1.0),
                            ema_decay=getattr(ema_cfg, "ema_decay", 0.0) if use_ema else 0.0)
            if old is not None and len(getattr(old, "state", {})) > 0 and hasattr(old, "state_dict"):
                try:  # a resumed run: run() already loaded optimizer.pt into the torch optimizer (:2189-2222)
                    opt.load_torch_state_dict(old.state_dict())
                    # moments and step come from the file; the hyper-parameters of the CURRENT config win (:2215-2218)
                    opt.param_groups[0].update(lr=lr, betas=tuple(op.get("betas", (0.9, 0.999))), eps=op.get("eps", 1e-6),
                                               weight_decay=op.get("weight_decay", 1e-2))
                    opt._hyper_host = None
                    opt.sync_hyper()
                except Exception as e:  # as the reference does: log and continue with fresh moments
                    print(f"sd_trainer_b200: could not carry the optimizer state over: {e}")
            self.optimizer = opt
            params = getattr(tc, "lr_scheduler_params", None) or {}
            if get_lr_scheduler is not None:
                params = dict(params)
                if "max_iterations" not in params:
                    params["total_iters"] = tc.steps
                self.lr_scheduler = get_lr_scheduler(getattr(tc, "lr_scheduler", "constant"), opt, **params)
            if self.ema is not None or use_ema:
                self.ema = FusedEMA(opt)
                if hasattr(self.sd, "ema"):
                    self.sd.ema = self.ema
            self._b200_steps = {}
            self._b200_mid_accumulation = False

        def _b200_check_supported(self, batch):
            tc = self.train_config
            if not getattr(self.sd, "is_flux", False) and not _is_unet(getattr(self.sd, "unet", None)):
                raise NotImplementedError("sd_trainer_b200: the fused step covers FLUX and the SD1.5 / SDXL UNet (BASELINE.json "
                                          "configs[0..2]); other architectures run their adapters through LoRAModule.forward "
                                          "under the stock trainer")
            for flag in ("do_prior_divergence", "train_turbo", "do_guided_loss", "inverted_mask_prior", "do_signal_amplification"):
                if getattr(tc, flag, False):
                    raise NotImplementedError(f"sd_trainer_b200: train.{flag} is outside the fused default path")
            if self._b200_preserving() and not getattr(self.sd, "is_flux", False):
                raise NotImplementedError("sd_trainer_b200: diff_output_preservation / blank_prompt_preservation are built for the "
                                          "FLUX step (prior-prediction target); not for the UNet step")
            if getattr(self, "adapter", None) is not None or getattr(self, "embedding", None) is not None:
                raise NotImplementedError("sd_trainer_b200: adapters / textual inversion are outside the fused default path")
            if str(getattr(tc, "loss_type", "mse")) != "mse":
                raise NotImplementedError(f"sd_trainer_b200: loss_type {tc.loss_type!r} (the fused loss kernel is 'mse')")

        def _b200_preserving(self):
            tc = self.train_config
            return bool(getattr(tc, "diff_output_preservation", False) or getattr(tc, "blank_prompt_preservation", False))

        def _b200_preservation_embeds(self, batch, conditioned_prompts, n):
            """The embeddings of the preservation pass (SDTrainer.py:1697-1705, 1772-1790, 2184-2193): the class prompt (trigger
            word replaced, per item) for diff_output_preservation, the blank prompt for blank_prompt_preservation."""
            tc = self.train_config
            if getattr(tc, "diff_output_preservation", False):
                pe = getattr(batch, "dop_prompt_embeds", None)  # cached to disk with the trigger word replaced per dataset
                if pe is None and getattr(self, "cached_dop_class_embeds", None) is not None:
                    pe = self.cached_dop_class_embeds           # no per-item cache: the class-only embeds
                if pe is None:
                    items = getattr(batch, "file_items", None) or [None] * len(conditioned_prompts)

                    def swap(prompt, item):
                        trig = getattr(item, "trigger_word", None) or getattr(self, "trigger_word", None)
                        return prompt if trig is None else prompt.replace(trig, tc.diff_output_preservation_class)

                    with torch.no_grad():
                        pe = self.sd.encode_prompt([swap(p, it) for p, it in zip(conditioned_prompts, items)],
                                                   long_prompts=getattr(self, "do_long_prompts", False))
                return pe.expand_to_batch(n) if hasattr(pe, "expand_to_batch") else pe
            blank = getattr(self, "cached_blank_embeds", None)
            if blank is None:
                with torch.no_grad():
                    blank = self.sd.encode_prompt("")
                self.cached_blank_embeds = blank
            return blank.expand_to_batch(n) if hasattr(blank, "expand_to_batch") else blank

        def _b200_step_for(self, latents, text_embeds, preservation=False):
            from .train_step import FluxLoRATrainStep

            key = (tuple(latents.shape), int(text_embeds.shape[1]), bool(preservation))
            step = self._b200_steps.get(key)
            if step is None and preservation:
                # the preservation term (SDTrainer.py:2182-2219): MSE(active prediction, inactive-network prediction) on the
                # preservation embeddings x multiplier = the step's prior-prediction target mode, as a further micro-batch
                tc = self.train_config
                mult = tc.diff_output_preservation_multiplier if getattr(tc, "diff_output_preservation", False) \
                    else tc.blank_prompt_preservation_multiplier
                step = FluxLoRATrainStep(self.sd.unet, self.network, self.optimizer, batch_size=latents.shape[0],
                                         latent_shape=tuple(latents.shape[1:]), text_len=int(text_embeds.shape[1]),
                                         guidance_scale=float(getattr(tc, "cfg_scale", 1.0)), use_cuda_graph=True,
                                         prior_target=True, loss_multiplier=float(mult))
                self._b200_steps[key] = step
            if step is None and not getattr(self.sd, "is_flux", False):
                from .unet import UNetLoRATrainStep

                tc = self.train_config
                step = UNetLoRATrainStep(self.sd.unet, self.network, self.optimizer,
                                         prediction_type=str(getattr(self.sd, "prediction_type", "epsilon")),
                                         min_snr_gamma=getattr(tc, "min_snr_gamma", None), snr_gamma=getattr(tc, "snr_gamma", None),
                                         use_cuda_graph=True)
                self._b200_steps[key] = step
            if step is None:
                tc = self.train_config
                step = FluxLoRATrainStep(self.sd.unet, self.network, self.optimizer, batch_size=latents.shape[0],
                                         latent_shape=tuple(latents.shape[1:]), text_len=int(text_embeds.shape[1]),
                                         guidance_scale=float(getattr(tc, "cfg_scale", 1.0)), use_cuda_graph=True)
                self._b200_steps[key] = step
            return step

        # ---------------------------------------------------------------- the step
        def hook_train_loop(self, batch):
            batch_list = batch if isinstance(batch, list) else [batch]
            total = 0.0
            device = self.device_torch
            n = len(batch_list)
            for i, b in enumerate(batch_list):
                self._b200_check_supported(b)
                b = self.preprocess_batch(b)
                with torch.no_grad():
                    noisy_latents, noise, timesteps, conditioned_prompts, imgs = self.process_general_training_batch(b)
                pe = getattr(b, "prompt_embeds", None)
                if pe is None:  # not cached: the frozen text encoders, as SDTrainer.py:1750 does
                    with torch.no_grad():
                        pe = self.sd.encode_prompt(conditioned_prompts, long_prompts=getattr(self, "do_long_prompts", False))
                text, pooled = pe.text_embeds, pe.pooled_embeds
                nw = getattr(b, "get_network_weight_list", None)
                if nw is not None:  # per-sample adapter strength (SDTrainer.py:1558)
                    self.network.multiplier = nw()
                step = self._b200_step_for(b.latents, text)
                lat_d, noise_d, ts_d = b.latents.to(device), noise.to(device), timesteps.to(device).float()
                step.load_batch(lat_d, noise_d, ts_d, text.to(device), pooled.to(device) if pooled is not None else None)
                first = (i == 0) and not self._b200_mid_accumulation
                last = (i == n - 1) and not getattr(self, "is_grad_accumulation_step", False)
                preserving = self._b200_preserving()
                loss = step.run(first_micro_batch=first, last_micro_batch=last and not preserving)
                step.loss_host.copy_(loss, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                normal = float(step.loss_host[0])
                total += normal
                if preserving:  # second micro-batch of the same step: gradients add up, one optimizer step (loss = normal + preservation)
                    ppe = self._b200_preservation_embeds(b, conditioned_prompts, lat_d.shape[0])
                    pstep = self._b200_step_for(b.latents, ppe.text_embeds, preservation=True)
                    pstep.load_batch(lat_d, noise_d, ts_d, ppe.text_embeds.to(device), ppe.pooled_embeds.to(device))
                    ploss = pstep.run(first_micro_batch=False, last_micro_batch=last)
                    pstep.loss_host.copy_(ploss, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                    pres = float(pstep.loss_host[0]) * float(getattr(pstep, "loss_multiplier", 1.0))  # (the kernel scales dpred only)
                    total += pres
                    logs = getattr(self, "additional_logs", None)
                    if isinstance(logs, dict):
                        logs["loss/normal"], logs["loss/preservation"] = normal, pres
            self._b200_mid_accumulation = bool(getattr(self, "is_grad_accumulation_step", False))
            # clip_grad_norm_, optimizer.step and ema.update already happened inside step.run (one fused launch sequence)
            self.lr_scheduler.step()
            loss_dict = OrderedDict({"loss": total / n})
            if hasattr(self, "end_of_training_loop"):
                self.end_of_training_loop()
            return loss_dict

    SDTrainerB200.__name__ = "SDTrainerB200"
    return SDTrainerB200
