"""One optimizer step of FLUX LoRA training on the B200 path — the body of
`SDTrainer.hook_train_loop` (extensions_built_in/sd_trainer/SDTrainer.py:2243-2317) for the default
flow-matching configuration, with cached latents and cached text embeddings:

    zero_grad -> add_noise (custom_flowmatch_sampler.py:91-102) + pack (stable_diffusion_model.py:2157-2172)
              -> predict_noise through the LoRA-wrapped DiT (`with network:`)
              -> target = noise - latents, MSE (SDTrainer.py:644-646, 916, 987-1013) -> backward
              -> [N GPUs] all-reduce(avg) of the flat LoRA gradient buffer over NCCL
              -> clip_grad_norm_(max_grad_norm) -> AdamW -> EMA -> {'loss': float}

Everything between the host->device copy of the batch and the device->host read of the loss is a static
launch schedule, captured once into two CUDA graphs (forward+backward, optimizer) with the gradient
all-reduce between them.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import calc_loss, ops
from . import timesteps as ts


def make_img_ids(h_lat: int, w_lat: int, device) -> torch.Tensor:
    """img_ids[..., 1] = row, [..., 2] = col over the packed (h/2, w/2) grid (stable_diffusion_model.py:2174-2178)."""
    ids = torch.zeros(h_lat // 2, w_lat // 2, 3, device=device)
    ids[..., 1] = ids[..., 1] + torch.arange(h_lat // 2, device=device)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w_lat // 2, device=device)[None, :]
    return ids.reshape(-1, 3)


class FluxLoRATrainStep:
    def __init__(self, model, network, optimizer, *, batch_size, latent_shape=(16, 128, 128), text_len=512,
                 guidance_scale=1.0, loss_multiplier=1.0, use_cuda_graph=True, process_group=None,
                 timestep_type="sigmoid", num_train_timesteps=1000, min_denoising_steps=0, max_denoising_steps=999,
                 linear_timesteps=False, linear_timesteps2=False, noise_multiplier=1.0, latent_multiplier=1.0,
                 use_loss_options=False, prior_target=False):
        self.model, self.network, self.optimizer = model, network, optimizer
        dev = model.device
        self.dev = dev
        C, H, W = latent_shape
        B = batch_size
        cfg = model.cfg
        self.B, self.C, self.H, self.W, self.Lt = B, C, H, W, text_len
        self.loss_multiplier = float(loss_multiplier)
        self.use_cuda_graph = use_cuda_graph
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        bf = torch.bfloat16
        # static device-side batch buffers (targets of the per-step host->device copies)
        self.latents = torch.zeros((B, C, H, W), device=dev, dtype=bf)
        self.noise = torch.zeros((B, C, H, W), device=dev, dtype=bf)
        self.timesteps = torch.zeros(B, device=dev, dtype=torch.float32)  # 0..1000, as scheduler.timesteps[idx]
        self.text = torch.zeros((B, text_len, cfg.joint_attention_dim), device=dev, dtype=bf)
        self.pooled = torch.zeros((B, cfg.pooled_projection_dim), device=dev, dtype=bf)
        self.guidance = torch.full((B,), float(guidance_scale), device=dev, dtype=torch.float32)
        self.txt_ids = torch.zeros(text_len, 3, device=dev)
        self.img_ids = make_img_ids(H, W, dev)
        # batch-preparation config (TrainConfig defaults, toolkit/config_modules.py:377-417, :556)
        self.timestep_type, self.num_train_timesteps = timestep_type, int(num_train_timesteps)
        self.min_denoising_steps, self.max_denoising_steps = int(min_denoising_steps), int(max_denoising_steps)
        self.linear_timesteps, self.linear_timesteps2 = bool(linear_timesteps), bool(linear_timesteps2)
        self.noise_multiplier, self.latent_multiplier = float(noise_multiplier), float(latent_multiplier)
        # calculate_loss options (timestep weights / per-sample loss multipliers / mask): device vectors the fused loss
        # kernel reads every step, so a captured CUDA graph picks up new values
        self.use_loss_options = bool(use_loss_options or linear_timesteps or linear_timesteps2 or timestep_type == "weighted")
        self.sample_weight = torch.ones(B, device=dev, dtype=torch.float32) if self.use_loss_options else None
        self.mask = None
        self._table = None
        # prior prediction as the target (SDTrainer.get_prior_prediction :1211-1339 -> calculate_loss `target = prior_pred`
        # :619-621): the frozen model's own prediction on the same noisy latents (network inactive, no gradient)
        self.prior_target = bool(prior_target)
        self.loss_ws = torch.zeros(B + 1, device=dev, dtype=torch.float32)
        self.loss_host = torch.zeros(1, dtype=torch.float32)
        if torch.device(dev).type == "cuda":
            self.loss_host = self.loss_host.pin_memory()
        if self.world > 1:
            self._setup_replicas()
        self._graph_fb = None
        self._graph_opt = None
        self._warm = 0

    def _setup_replicas(self):
        """N > 1: what `accelerator.prepare(network / optimizer)` (BaseSDTrainProcess.py:744-779) would give a DDP run.
        (1) every replica starts from rank 0's adapter values (DDP broadcasts module state at construction; lora_down is
        drawn from each process's own RNG); (2) the all-reduce SUMS, so the 1/W of the average goes into the clip/AdamW
        kernel (hyper[7]) -- set here, not left to the caller, because with the default 1.0 the clipping norm would be
        W times too large; (3) the EMA shadow follows the broadcast values."""
        net, opt = self.network, self.optimizer
        if net.flat_params is None:
            net._flatten()
        torch.distributed.broadcast(net.flat_params, src=torch.distributed.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                                    group=self.pg)
        net.mark_params_changed()
        if opt is not None:
            opt.grad_prescale = 1.0 / self.world
            if getattr(opt, "ema", None) is not None:
                opt.ema.copy_(net.flat_params)
            opt._hyper_host = None
            opt.sync_hyper()

    # -- the launch schedule -------------------------------------------------------------------------
    def _forward_backward(self):
        net = self.network
        packed = ops.flow_add_noise(self.latents, self.noise, self.timesteps, pack=True)
        # the trainer passes timestep / 1000 (stable_diffusion_model.py:2196) and the bf16 model re-scales by 1000 in
        # bf16: both happen inside the timestep-embedding kernel (t_div=1000)
        prior = None
        if self.prior_target:
            net.is_active = False  # get_prior_prediction: `self.network.is_active = False`, torch.no_grad()
            prior = self.model.engine.forward(packed, self.timesteps, self.text, self.pooled, self.guidance, self.txt_ids,
                                              self.img_ids, save=False, t_div=1000.0)
        net.is_active = True  # `with network:` (SDTrainer.py:2229-2238 keeps it active through backward)
        try:
            eng = self.model.engine
            pred = eng.forward(packed, self.timesteps, self.text, self.pooled, self.guidance, self.txt_ids, self.img_ids,
                               save=True, t_div=1000.0)
            if prior is not None:
                _, _, dpred = ops.train_loss(pred.view(self.B, -1, pred.shape[-1]), self.latents, self.noise, target=prior,
                                             target_in_pred_layout=True, sample_weight=self.sample_weight, mask=self.mask,
                                             pack=True, gscale=self.loss_multiplier, loss_ws=self.loss_ws)
            elif self.use_loss_options or self.mask is not None:
                _, _, dpred = ops.train_loss(pred.view(self.B, -1, pred.shape[-1]), self.latents, self.noise,
                                             sample_weight=self.sample_weight, mask=self.mask, pack=True,
                                             gscale=self.loss_multiplier, loss_ws=self.loss_ws)
            else:
                _, _, dpred = ops.flow_loss(pred.view(self.B, -1, pred.shape[-1]), self.latents, self.noise, pack=True,
                                            gscale=self.loss_multiplier, loss_ws=self.loss_ws)
            eng.backward(dpred.view(-1, pred.shape[-1]))
        finally:
            net.is_active = False

    def _optimizer(self):
        self.optimizer.step()

    def _all_reduce(self):
        if self.world > 1:
            # sum over ranks; the 1/world of the average is folded into the clip/AdamW kernel (hyper[7])
            torch.distributed.all_reduce(self.network.flat_grads, op=torch.distributed.ReduceOp.SUM, group=self.pg)

    def load_batch(self, latents, noise, timesteps, text_embeds, pooled_embeds):
        """Host (pinned) or device tensors -> the static device buffers, asynchronously on the current stream."""
        self.latents.copy_(latents, non_blocking=True)
        self.noise.copy_(noise, non_blocking=True)
        self.timesteps.copy_(timesteps, non_blocking=True)
        self.text.copy_(text_embeds, non_blocking=True)
        self.pooled.copy_(pooled_embeds, non_blocking=True)

    # -- batch preparation: BaseSDTrainProcess.process_general_training_batch for cached latents (:1036-1478) ---------
    def prepare_batch(self, latents, text_embeds, pooled_embeds, generator=None, loss_multiplier=None, mask=None):
        """latents [B, C, H, W] -> loads the step's device buffers with what the reference's batch preparation yields:
        the per-step timestep table (`set_train_timesteps`, :1195-1229), index draw (`content_or_style == 'balanced'`,
        :1301-1318), `timesteps = table[idx]` (:1323), noise = randn_like x noise_multiplier (:1327, :1351), latents x
        latent_multiplier (:1402); add_noise itself (:1421) is the first kernel of the step.  Timestep weights / loss
        multipliers go to the device vector of the fused loss.  `generator`: a torch.Generator on the step's device (the
        RNG stream is torch's).  Returns (timesteps, indices) for logging / tests."""
        dev = self.dev
        lat = latents.to(dev, torch.bfloat16)
        if self.latent_multiplier != 1.0:
            lat = lat * self.latent_multiplier
        table = ts.set_train_timesteps(self.num_train_timesteps, dev, self.timestep_type, generator=generator, latents=lat,
                                       patch_size=1)
        idx = ts.sample_timestep_indices(self.B, dev, self.min_denoising_steps, self.max_denoising_steps, flowmatch=True,
                                         generator=generator)
        timesteps = ts.timesteps_for_batch(table, idx).float()
        noise = torch.randn(lat.shape, device=dev, dtype=lat.dtype, generator=generator)
        if self.noise_multiplier != 1.0:
            noise = noise * self.noise_multiplier
        self._table = table
        self.load_batch(lat, noise, timesteps, text_embeds, pooled_embeds)
        self.set_loss_options(timesteps, loss_multiplier=loss_multiplier, mask=mask)
        return timesteps, idx

    def set_loss_options(self, timesteps, loss_multiplier=None, mask=None):
        """Per-sample weights (timestep weights x loss multipliers, SDTrainer.py:923-943, :994) and the mask multiplier
        (:953-959) of the step's loss; needs `use_loss_options=True` at construction when a CUDA graph is in use."""
        want = self.linear_timesteps or self.linear_timesteps2 or self.timestep_type == "weighted" or loss_multiplier is not None
        if want:
            if self.sample_weight is None:
                raise RuntimeError("construct FluxLoRATrainStep(use_loss_options=True) to use timestep weights / loss multipliers")
            v = calc_loss.loss_vectors(timesteps, is_flow_matching=True, flow_table=self._table,
                                       linear_timesteps=self.linear_timesteps, linear_timesteps2=self.linear_timesteps2,
                                       timestep_type=self.timestep_type, loss_multiplier=loss_multiplier, device=self.dev)
            if v["sample_weight"] is None:
                self.sample_weight.fill_(1.0)
            else:
                self.sample_weight.copy_(v["sample_weight"], non_blocking=True)
        if mask is not None:
            m = mask.to(self.dev, torch.float32).contiguous()
            if self.mask is None:
                if self._graph_fb is not None:
                    raise RuntimeError("a mask must be present before the step is captured into its CUDA graph")
                self.mask = torch.empty_like(m)
            self.mask.copy_(m, non_blocking=True)

    # -- deterministic evaluation: BaseSDTrainProcess.validate (:1681-1743) --------------------------------------------
    @torch.no_grad()
    def validate(self, latents_list, embeds_list, sigmas=(1.0, 0.75, 0.5, 0.25)):
        """Validation loss with the reference's protocol: fixed CPU-generator noise per image (seeds 42 + i, :1666-1672),
        all sigmas as one batch (timesteps = sigma x 1000), network active at multiplier 1.0, no gradient;
        `mse_loss(pred.float(), (noise - latents).float())` per image, mean over images.  -> python float."""
        net = self.network
        start_multiplier = net.multiplier
        net.multiplier = 1.0
        n = len(sigmas)
        t = torch.tensor([s * 1000.0 for s in sigmas], device=self.dev, dtype=torch.float32)
        losses = []
        was_active = net.is_active
        net.is_active = True
        try:
            for i, (lat, (text, pooled)) in enumerate(zip(latents_list, embeds_list)):
                g = torch.Generator(device="cpu").manual_seed(42 + i)
                noise = torch.randn(lat.shape, generator=g, dtype=torch.float32)
                lat_b = torch.cat([lat.to(self.dev, torch.bfloat16)] * n, 0)
                noise_b = torch.cat([noise.to(self.dev, torch.bfloat16)] * n, 0)
                text_b = torch.cat([text.to(self.dev, torch.bfloat16)] * n, 0)
                pooled_b = torch.cat([pooled.to(self.dev, torch.bfloat16)] * n, 0)
                packed = ops.flow_add_noise(lat_b, noise_b, t, pack=True)
                H, W = lat_b.shape[2], lat_b.shape[3]
                pred = self.model.engine.forward(packed, t, text_b, pooled_b, torch.full((n,), float(self.guidance[0]), device=self.dev),
                                                 torch.zeros(text_b.shape[1], 3, device=self.dev), make_img_ids(H, W, self.dev),
                                                 save=False, t_div=1000.0)
                tot, _, _ = ops.flow_loss(pred.view(n, -1, pred.shape[-1]), lat_b, noise_b, pack=True, want_grad=False)
                losses.append(tot.clone())
        finally:
            net.is_active = was_active
            net.multiplier = start_multiplier
        return float(torch.stack(losses).mean())

    # -- save / resume from the device flat buffers (BaseSDTrainProcess.save :505-721, run :2046-2060, :2189-2222) ----
    def save(self, save_root, name, step, epoch=0, dtype=torch.float16, **kw):
        from . import checkpoint

        torch.cuda.synchronize()
        return checkpoint.save_checkpoint(self.network, self.optimizer, save_root, name, step=step, epoch=epoch, dtype=dtype, **kw)

    def resume(self, save_root, name, **kw):
        """-> (path, step, epoch).  Weights, EMA shadow, AdamW moments and step counter are restored into the flat device
        buffers; the bf16 operand packs are refreshed by the next step (captured graphs stay valid: same buffers)."""
        from . import checkpoint

        out = checkpoint.resume(self.network, self.optimizer, save_root, name, **kw)
        self.network.mark_params_changed()
        self.network.refresh_packs(force=True)
        return out

    def _load_dict(self, b):
        self.load_batch(b["latents"], b["noise"], b["timesteps"], b["text_embeds"], b["pooled_embeds"])

    def run(self, first_micro_batch=True, last_micro_batch=True):
        """Launch one (micro-)step on the resident batch; returns the device loss scalar (no host sync).

        Gradient accumulation (`SDTrainer.hook_train_loop`, SDTrainer.py:2250-2268: gradients of the micro-batches are
        SUMMED, the optimizer runs after the last one): `first_micro_batch` zeroes the flat gradient buffer,
        `last_micro_batch` runs all-reduce + clip + AdamW + EMA."""
        self.optimizer.sync_hyper()
        if first_micro_batch:
            self.network.ensure_grad_views()
            self.network.flat_grads.zero_()
        if not self.use_cuda_graph or self._warm < 2:
            self._forward_backward()
            if last_micro_batch:
                self._all_reduce()
                self._optimizer()
            self._warm += 1
            return self.loss_ws[self.B:self.B + 1]
        if self._graph_fb is None:
            torch.cuda.synchronize()
            self._graph_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_fb):
                self._forward_backward()
            self._graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_opt):
                self._optimizer()
            # capture does not execute: run the step that was just recorded
        self._graph_fb.replay()
        if last_micro_batch:
            self._all_reduce()
            self._graph_opt.replay()
        return self.loss_ws[self.B:self.B + 1]

    def hook_train_loop(self, batch) -> OrderedDict:
        """`batch` = dict(latents, noise, timesteps, text_embeds, pooled_embeds), or a LIST of such dicts (gradient
        accumulation, as the reference's `batch_list`); returns OrderedDict(loss=float) like SDTrainer.hook_train_loop
        (:2312-2317): the mean of the micro-batch losses.  The `.item()` there is the same one device->host sync here."""
        batches = batch if isinstance(batch, (list, tuple)) else [batch]
        total = 0.0
        for i, b in enumerate(batches):
            self._load_dict(b)
            loss = self.run(first_micro_batch=(i == 0), last_micro_batch=(i == len(batches) - 1))
            self.loss_host.copy_(loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            total += float(self.loss_host[0])
        return OrderedDict(loss=total / len(batches))


class WanLoRATrainStep(FluxLoRATrainStep):
    """The same step for the Wan2.1 video DiT (BASELINE.json configs[3]): `Wan21.get_noise_prediction`
    (toolkit/models/wan21/wan21.py:578-603: raw 0..1000 timestep, UMT5 embeddings, no pooled vector / guidance) and
    `get_loss_target` = noise - latents (:717-724); flow-matching add_noise with the scheduler of :80-84 (shift 3.0 is a
    property of the timestep TABLE, i.e. of `prepare_batch(timestep_type='shift')`, not of add_noise).  Latents are 5-D
    [B, 16, F, H, W]; the (1, 2, 2) patchify is the packed layout of the add-noise kernel on the [B, 16, F*H, W] view."""

    def __init__(self, model, network, optimizer, *, batch_size, latent_shape=(16, 13, 64, 64), text_len=512,
                 loss_multiplier=1.0, use_cuda_graph=True, process_group=None, **kw):
        self.model, self.network, self.optimizer = model, network, optimizer
        dev = model.device
        self.dev = dev
        C, Fr, H, W = latent_shape
        B = batch_size
        self.B, self.C, self.F, self.H, self.W, self.Lt = B, C, Fr, H, W, text_len
        self.grid = (Fr, H // 2, W // 2)
        self.loss_multiplier = float(loss_multiplier)
        self.use_cuda_graph = use_cuda_graph
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        bf = torch.bfloat16
        self.latents = torch.zeros((B, C, Fr * H, W), device=dev, dtype=bf)  # 4-D view of the 5-D latents (F and H merged)
        self.noise = torch.zeros((B, C, Fr * H, W), device=dev, dtype=bf)
        self.timesteps = torch.zeros(B, device=dev, dtype=torch.float32)
        self.text = torch.zeros((B, text_len, model.cfg.text_dim), device=dev, dtype=bf)
        self.pooled = None
        self.timestep_type = kw.get("timestep_type", "shift")
        self.num_train_timesteps = int(kw.get("num_train_timesteps", 1000))
        self.min_denoising_steps, self.max_denoising_steps = int(kw.get("min_denoising_steps", 0)), int(kw.get("max_denoising_steps", 999))
        self.linear_timesteps = self.linear_timesteps2 = False
        self.noise_multiplier, self.latent_multiplier = float(kw.get("noise_multiplier", 1.0)), float(kw.get("latent_multiplier", 1.0))
        self.use_loss_options = bool(kw.get("use_loss_options", False))
        self.sample_weight = torch.ones(B, device=dev, dtype=torch.float32) if self.use_loss_options else None
        self.mask = None
        self._table = None
        self.loss_ws = torch.zeros(B + 1, device=dev, dtype=torch.float32)
        self.loss_host = torch.zeros(1, dtype=torch.float32)
        if torch.device(dev).type == "cuda":
            self.loss_host = self.loss_host.pin_memory()
        if self.world > 1:
            self._setup_replicas()
        self._graph_fb = None
        self._graph_opt = None
        self._warm = 0

    def load_batch(self, latents, noise, timesteps, text_embeds, pooled_embeds=None):
        B = self.B
        self.latents.copy_(latents.reshape(B, self.C, self.F * self.H, self.W), non_blocking=True)
        self.noise.copy_(noise.reshape(B, self.C, self.F * self.H, self.W), non_blocking=True)
        self.timesteps.copy_(timesteps, non_blocking=True)
        self.text.copy_(text_embeds, non_blocking=True)

    def _load_dict(self, b):
        self.load_batch(b["latents"], b["noise"], b["timesteps"], b["text_embeds"])

    def _forward_backward(self):
        net = self.network
        packed = ops.flow_add_noise(self.latents, self.noise, self.timesteps, pack=True)  # add_noise + (1, 2, 2) patchify
        net.is_active = True
        try:
            eng = self.model.engine
            pred = eng.forward(packed, self.timesteps, self.text, self.grid, save=True)
            if self.use_loss_options or self.mask is not None:
                _, _, dpred = ops.train_loss(pred.view(self.B, -1, pred.shape[-1]), self.latents, self.noise,
                                             sample_weight=self.sample_weight, mask=self.mask, pack=True,
                                             gscale=self.loss_multiplier, loss_ws=self.loss_ws)
            else:
                _, _, dpred = ops.flow_loss(pred.view(self.B, -1, pred.shape[-1]), self.latents, self.noise, pack=True,
                                            gscale=self.loss_multiplier, loss_ws=self.loss_ws)
            eng.backward(dpred.view(-1, pred.shape[-1]))
        finally:
            net.is_active = False

    def validate(self, *a, **k):
        raise NotImplementedError("validate() is implemented for the FLUX step")
