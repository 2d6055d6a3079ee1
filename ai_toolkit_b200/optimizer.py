"""`B200AdamW`: grad-norm -> clip -> AdamW (-> EMA) over the network's flat fp32 buffers, 4 launches per step.

Replaces `accelerator.clip_grad_norm_(params, max_grad_norm)` + `torch.optim.AdamW(eps=1e-6).step()` +
`ExponentialMovingAverage.update()` (extensions_built_in/sd_trainer/SDTrainer.py:2278-2297,
toolkit/optimizer.py:78-79, toolkit/ema.py:100-152).  It is a `torch.optim.Optimizer` so that schedulers,
`state_dict()` / `load_state_dict()` (the reference saves `optimizer.pt`, BaseSDTrainProcess.py:702-714) and
`zero_grad()` keep working; hyper-parameters live in a small device buffer so that a captured CUDA graph of
the step picks up learning-rate changes.
"""
from __future__ import annotations

import torch

from . import ops


class B200AdamW(torch.optim.Optimizer):
    def __init__(self, network, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0,
                 ema_decay=0.0, grad_prescale=1.0, ema_use_num_updates=False):
        if network.flat_params is None:
            network._flatten()
        self.network = network
        params = [p for m in network.get_all_modules() for p in (m.lora_down.weight, m.lora_up.weight)]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        flat = network.flat_params
        dev = flat.device
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.ema = flat.clone() if ema_decay and ema_decay > 0 else None
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.ema_decay = float(ema_decay or 0.0)
        self.grad_prescale = float(grad_prescale)
        self.hyper = torch.zeros(8, device=dev, dtype=torch.float32)
        self.state_buf = torch.zeros(8, device=dev, dtype=torch.int64)  # 64 bytes: step counter + derived scalars
        # ExponentialMovingAverage(use_num_updates=...): the trainer leaves it False = constant decay
        # (BaseSDTrainProcess.py:798-803); True = warm-up min(decay, (1 + n) / (10 + n)) (toolkit/ema.py:121-128)
        self.ema_use_num_updates = bool(ema_use_num_updates)
        self.state_buf[1] = int(self.ema_use_num_updates)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self._hyper_host = None
        self.sync_hyper()

    def sync_hyper(self):
        g = self.param_groups[0]
        vals = [g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.max_grad_norm, self.ema_decay,
                self.grad_prescale]
        if vals != self._hyper_host:
            self.hyper.copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper_host = vals

    @torch.no_grad()
    def reset_ema(self):
        """EMA shadow := current parameters (the reference constructs its ExponentialMovingAverage after load_weights,
        BaseSDTrainProcess.py:2053 then :2229)."""
        if self.ema is not None:
            self.ema.copy_(self.network.flat_params)

    @torch.no_grad()
    def zero_grad(self, set_to_none: bool = False):
        self.network.ensure_grad_views()
        self.network.flat_grads.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        net = self.network
        if net.flat_params.device.type != "cuda":
            raise RuntimeError("B200AdamW runs on the B200 only (no CPU fallback)")
        if not torch.cuda.is_current_stream_capturing():  # (a captured step is refreshed by FluxLoRATrainStep.run)
            self.sync_hyper()  # pick up what an lr scheduler wrote into param_groups (no copy when nothing changed)
        ops.grad_sumsq(net.flat_grads, self.sumsq)
        ops.clip_adamw(net.flat_params, net.flat_grads, self.exp_avg, self.exp_avg_sq, self.sumsq, self.hyper, self.state_buf,
                       ema=self.ema, norm_out=self.grad_norm)
        ops.repack_lora(net.flat_params, net.pack_buf, net._repack_table, net._repack_entries)
        net._pack_dirty = False
        return None

    # -- interchange with torch.optim.AdamW (the reference saves `optimizer.pt`, BaseSDTrainProcess.py:702-714) ------
    def torch_state_dict(self):
        """State in torch.optim.AdamW's own layout (per-parameter `step` / `exp_avg` / `exp_avg_sq`, one param group),
        so that a run can be resumed by the reference's optimizer and vice versa."""
        step = float(self.state_buf[0].item())
        state, off = {}, 0
        params = self.param_groups[0]["params"]
        for i, p in enumerate(params):
            n = p.numel()
            state[i] = {"step": torch.tensor(step), "exp_avg": self.exp_avg[off:off + n].view(p.shape).detach().cpu().clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).detach().cpu().clone()}
            off += n
        g = self.param_groups[0]
        group = {"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"], "weight_decay": g["weight_decay"], "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def load_torch_state_dict(self, sd):
        params = self.param_groups[0]["params"]
        off, step = 0, 0
        for i, p in enumerate(params):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = int(float(st["step"]))
            off += n
        self.state_buf.zero_()
        self.state_buf[0] = step
        self.state_buf[1] = int(self.ema_use_num_updates)
        g = sd["param_groups"][0]
        self.param_groups[0].update({k: g[k] for k in ("lr", "betas", "eps", "weight_decay") if k in g})
        self._hyper_host = None
        self.sync_hyper()

    # -- checkpointing (flat tensors instead of per-parameter dicts) ---------------------------------
    def state_dict(self):
        return {"b200_flat": True, "exp_avg": self.exp_avg.cpu(), "exp_avg_sq": self.exp_avg_sq.cpu(),
                "ema": None if self.ema is None else self.ema.cpu(), "step": int(self.state_buf[0].item()),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        if not sd.get("b200_flat"):
            raise ValueError("not a B200AdamW state dict")
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        if self.ema is not None and sd.get("ema") is not None:
            self.ema.copy_(sd["ema"])
        self.state_buf.zero_()
        self.state_buf[0] = int(sd["step"])
        self.state_buf[1] = int(self.ema_use_num_updates)
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update(saved)
        self._hyper_host = None
        self.sync_hyper()
