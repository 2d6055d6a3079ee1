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

from . import ops


def make_img_ids(h_lat: int, w_lat: int, device) -> torch.Tensor:
    """img_ids[..., 1] = row, [..., 2] = col over the packed (h/2, w/2) grid (stable_diffusion_model.py:2174-2178)."""
    ids = torch.zeros(h_lat // 2, w_lat // 2, 3, device=device)
    ids[..., 1] = ids[..., 1] + torch.arange(h_lat // 2, device=device)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w_lat // 2, device=device)[None, :]
    return ids.reshape(-1, 3)


class FluxLoRATrainStep:
    def __init__(self, model, network, optimizer, *, batch_size, latent_shape=(16, 128, 128), text_len=512,
                 guidance_scale=1.0, loss_multiplier=1.0, use_cuda_graph=True, process_group=None):
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
        net.is_active = True  # `with network:` (SDTrainer.py:2229-2238 keeps it active through backward)
        try:
            eng = self.model.engine
            pred = eng.forward(packed, self.timesteps, self.text, self.pooled, self.guidance, self.txt_ids, self.img_ids,
                               save=True, t_div=1000.0)
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
            self.load_batch(b["latents"], b["noise"], b["timesteps"], b["text_embeds"], b["pooled_embeds"])
            loss = self.run(first_micro_batch=(i == 0), last_micro_batch=(i == len(batches) - 1))
            self.loss_host.copy_(loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            total += float(self.loss_host[0])
        return OrderedDict(loss=total / len(batches))
