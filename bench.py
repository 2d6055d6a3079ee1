#!/usr/bin/env python
"""bench.py — train-steps/sec of FLUX.1-dev LoRA (r=16, bs=1/GPU, 1024^2) on the B200-native path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = SDTrainer.hook_train_loop: add_noise + pack -> forward through the
LoRA-wrapped frozen DiT -> flow-matching MSE -> backward (dX through 11.9 B frozen weights, dA/dB for 494 adapters)
-> [N GPUs: all-reduce of the flat LoRA gradient over NCCL] -> clip_grad_norm_(1.0) -> AdamW(eps 1e-6) -> EMA.
Synthetic latents / embeddings / weights of the true shapes (no network for checkpoints).  Prints ONE JSON line.

  value     steps/s of the whole job with the batch resident in HBM (CUDA events, max over ranks)
  e2e       the same through FluxLoRATrainStep.hook_train_loop with pinned HOST batches: H2D copies and the D2H
            read of the loss inside the timed region
  roofline  the dominant kernel (fused LoRA-Linear tcgen05 GEMM) timed live, against MEASURED_PEAKS.json
  cpu_baseline  the oracle (eager PyTorch restatement of the reference path) on the host cores, bounded sample

`--impl reference` times the reference's CPU implementation of the path: ostris/ai-toolkit is not pip-installable
(no setup.py / pyproject) and its model code is diffusers (absent), so this is the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RANK = 16
LATENT = (16, 128, 128)  # 1024^2 image -> 16 x 128 x 128 latent -> 4096 packed tokens
TEXT_LEN = 512
METRIC = "train-steps/sec FLUX.1-dev LoRA r=16 bs=1 1024^2"


def workload_name(args):
    if getattr(args, "model", "flux") == "wan":
        return f"Wan2.1-T2V-1.3B LoRA r={args.rank} bs={args.batch}/GPU 49x512x512 clip (BASELINE.json configs[3])"
    if args.batch == 1 and args.rank == 16:
        return "FLUX.1-dev LoRA r=16 bs=1/GPU 1024^2 (BASELINE.json configs[2])"
    return f"FLUX.1-dev LoRA r={args.rank} bs={args.batch}/GPU 1024^2 (BASELINE.json configs[4] rank sweep)"


def flux_flops(B=1, r=RANK, I=4096, T=512, D=3072, M=12288, n_double=19, n_single=38):
    """Algorithmic FLOPs of one step: F = 2 F_lin + 3.5 F_attn + 3 F_lora."""
    L = I + T
    f_lin = n_double * (2 * B * (I + T) * (4 * D * D + 2 * D * M) + 2 * B * 2 * 6 * D * D) \
        + n_single * (2 * B * L * (3 * D * D + D * M + (D + M) * D) + 2 * B * 3 * D * D)
    f_attn = (n_double + n_single) * 4 * B * 24 * L * L * 128
    lora = 0
    for toks, i, o, cnt in ((I, D, D, 4 * n_double), (T, D, D, 4 * n_double), (I, D, M, n_double), (I, M, D, n_double),
                            (T, D, M, n_double), (T, M, D, n_double), (1, D, 6 * D, 2 * n_double),
                            (L, D, D, 3 * n_single), (L, D, M, n_single), (L, D + M, D, n_single), (1, D, 3 * D, n_single)):
        lora += cnt * 2 * B * toks * r * (i + o)
    return 2 * f_lin + 3.5 * f_attn + 3 * lora, f_lin, f_attn, lora


WAN_LATENT = (16, 13, 64, 64)  # 49 frames x 512 x 512 -> (49 - 1) / 4 + 1 = 13 latent frames x 64 x 64 -> 13,312 tokens


def wan_flops(B=1, r=RANK, L=13312, T=512, D=1536, F=8960, H=12, layers=30):
    """Algorithmic FLOPs of one Wan2.1-T2V-1.3B LoRA step: 2 F_lin + 3.5 F_attn + 3 F_lora."""
    f_lin = layers * (2 * B * L * (6 * D * D + 2 * D * F) + 2 * B * T * 2 * D * D)
    f_attn = layers * (4 * B * H * L * L * 128 + 4 * B * H * L * T * 128)
    f_lora = layers * 2 * B * r * (L * (8 * D + 4 * D + 2 * (D + F)) + T * 4 * D)
    return 2 * f_lin + 3.5 * f_attn + 3 * f_lora, f_lin, f_attn, f_lora


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def host_cores() -> int:
    """Threads this process may really use: the scheduler affinity mask, bounded by the cgroup CPU quota (cpu.max) --
    `os.cpu_count()` reports the machine, not the lease, and oversubscribing a quota'd container makes CPU timings
    irreproducible (round-1 BENCH vs SCALE differed 32x)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    return max(1, n)


def profile_numbers(prefixes=("r2_gemm_fwd", "r1_gemm_fwd")):
    """dram bytes / tensor-pipe activity of the dominant kernel from the newest COMMITTED ncu --set full summary
    (profiles/*_ncu_full_summary.csv, written by tools/ncu_summary.py from the .ncu-rep of tools/r2_profile_trip.sh)."""
    import csv

    for pre in prefixes:
        path = os.path.join(ROOT, "profiles", f"{pre}_ncu_full_summary.csv")
        if not os.path.exists(path):
            continue
        try:
            rows = list(csv.reader(open(path)))
            hdr, units, row = rows[0], rows[1], rows[2]
            get = lambda k: float(row[hdr.index(k)])  # noqa: E731
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd = get("dram__bytes_read.sum") * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
            wr = get("dram__bytes_write.sum") * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
            return {"traffic": rd + wr, "tensor_pipe_active_pct_ncu": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "ncu_us": get("gpu__time_duration.sum"), "kernel_ncu": row[0][:80], "source": f"profiles/{os.path.basename(path)}"}
        except Exception as e:  # malformed summary: say so instead of inventing numbers
            return {"traffic": None, "source": f"profiles/{os.path.basename(path)} (unreadable: {e})"}
    return {"traffic": None, "source": None}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_double=1, n_single=1, steps=3, warmup=1, tokens_img=1024, tokens_txt=128, rank=None):
    """Eager oracle (oracle/flux_ref.py + oracle/lora_ref.py: the reference's LoRA forward restated, pinned bit-for-bit to
    the reference classes in tests/test_oracle_pinned.py) on the host CPU in float32 (BASELINE.json configs[0]: the
    reference's CPU path is float32), FLUX width (D = 3072, 24 heads, MLP 12288) but a bounded number of blocks / tokens.
    Returns (median seconds per sample-step, [all step seconds], F_sample, description, threads)."""
    import torch

    from oracle import flux_ref, lora_ref

    rank = rank or RANK
    threads = host_cores()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    cfg = flux_ref.FluxConfig(num_layers=n_double, num_single_layers=n_single)
    model = flux_ref.FluxTransformer2DModel(cfg).to(torch.float32).requires_grad_(False)
    net = lora_ref.LoRANetworkRef(model, lora_dim=rank)
    params = [p for l in net.loras for p in (l.lora_down.weight, l.lora_up.weight)]
    opt = torch.optim.AdamW(params, lr=1e-4, eps=1e-6)
    side = int(round((tokens_img * 4) ** 0.5))
    lat = torch.randn(1, 16, side, side)
    noise = torch.randn_like(lat)
    t = torch.tensor([500.0])
    text = torch.randn(1, tokens_txt, 4096) * 0.1
    pooled = torch.randn(1, 768)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        noisy = lora_ref.add_noise_flowmatch(lat, noise, t)
        with net:
            pred = lora_ref.flux_predict(model, noisy, t, text, pooled, 1.0, flux_ref.pack_latents, flux_ref.unpack_latents,
                                         flux_ref.make_img_ids)
            loss = lora_ref.flow_loss(pred, lat, noise)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        float(loss)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    I = (side // 2) ** 2
    f_sample = flux_flops(1, rank, I, tokens_txt, n_double=n_double, n_single=n_single)[0]
    med = sorted(times)[len(times) // 2]
    desc = (f"oracle port (eager PyTorch float32) on {threads} host threads (affinity / cgroup quota; machine has "
            f"{os.cpu_count()}): FLUX-width blocks {n_double} double + {n_single} single, {I}+{tokens_txt} tokens, r={rank}; "
            f"{warmup} warm-up + {steps} timed steps, median {med:.2f} s (min {min(times):.2f}, max {max(times):.2f}); steps/s "
            f"extrapolated linearly in F_step ({f_sample / 1e12:.2f} of {flux_flops(1, rank)[0] / 1e12:.1f} TFLOP)")
    return med, times, f_sample, desc, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if getattr(args, "model", "flux") != "flux":  # the CPU arm times the headline workload; the other models carry `gpu_reference`
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU reference arm covers --model flux (BASELINE.json's metric); "
                                                              f"--model {args.model} reports the eager reference-style step as gpu_reference"}))
        return
    sec, times, f_sample, desc, threads = cpu_reference_sample(steps=max(3, min(args.steps, 5)), warmup=max(1, min(args.warmup, 2)),
                                                               rank=args.rank)
    f_step = flux_flops(args.batch, args.rank)[0]
    value = 1.0 / (sec * f_step / f_sample)
    out = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": workload_name(args) + ", CPU sample extrapolated in F_step"},
           "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port", "sample": desc,
                            "step_seconds": times},
           "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def gpu_reference_leg(dev, steps=3, warmup=2, rank=None, batch=1):
    """the reference-style EAGER PyTorch step on the SAME B200 -- the denominator of the north
    star's ">= 4x the reference's 1xB200 PyTorch step time".  diffusers-named FLUX blocks (oracle/flux_ref.py) + the
    reference's LoRA forward as restated in oracle/lora_ref.py (fp32 side branch, network_mixins.py:304-342; pinned
    bit-for-bit to the reference class, which cannot travel to the GPU box) + torch SDPA + clip_grad_norm_ +
    torch.optim.AdamW(eps=1e-6); bf16 base, fp32 adapters; timed with gradient checkpointing (the reference default,
    toolkit/config_modules.py:413) and without (equal work to this repo's path, which stores activations)."""
    import torch

    from oracle import flux_ref, lora_ref

    rank = rank or RANK
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = flux_ref.FluxTransformer2DModel(flux_ref.flux_dev_config())
    torch.set_default_dtype(torch.float32)
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * 0.02)
    model.requires_grad_(False)
    net = lora_ref.LoRANetworkRef(model, lora_dim=rank).to(dev)
    with torch.no_grad():
        for l in net.loras:
            l.lora_up.weight.normal_(0, 0.02)
    params = [p for l in net.loras for p in (l.lora_down.weight, l.lora_up.weight)]
    opt = torch.optim.AdamW(params, lr=1e-4, eps=1e-6)
    lat = torch.randn(batch, *LATENT, device=dev).bfloat16()
    noise = torch.randn_like(lat)
    t = torch.full((batch,), 500.0, device=dev)
    text = (torch.randn(batch, TEXT_LEN, 4096, device=dev) * 0.1).bfloat16()
    pooled = torch.randn(batch, 768, device=dev).bfloat16()

    def step():
        opt.zero_grad(set_to_none=True)
        noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16)
        with net:
            pred = lora_ref.flux_predict(model, noisy, t, text, pooled, 1.0, flux_ref.pack_latents, flux_ref.unpack_latents,
                                         flux_ref.make_img_ids)
            loss = lora_ref.flow_loss(pred, lat, noise)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    out = {"what": "eager PyTorch reference-style step on the same GPU (oracle FLUX blocks + reference LoRA forward + torch "
                   "SDPA + torch AdamW), CUDA-event timed", "steps": steps, "warmup": warmup}
    for name, ck in (("checkpointing", True), ("no_checkpointing", False)):
        try:
            model.gradient_checkpointing = ck
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            out[name] = {"ms_per_step": e0.elapsed_time(e1) / steps, "loss": float(loss)}
        except Exception as e:  # e.g. out of memory without checkpointing at bs 4
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:100]}"}
            torch.cuda.empty_cache()
    out["peak_mem_gib"] = torch.cuda.max_memory_allocated() / 2 ** 30
    return out


def gpu_reference_leg_wan(dev, steps=3, warmup=2, rank=None, batch=1):
    """The same eager reference-style step for Wan2.1-T2V-1.3B (oracle/wan_ref.py blocks, whose attention is pinned to the
    reference's in-tree WanAttnProcessor2_0, + the reference LoRA forward + torch SDPA + torch AdamW)."""
    import torch

    from oracle import lora_ref, wan_ref

    rank = rank or RANK
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = wan_ref.WanTransformer3DModel(wan_ref.wan_1_3b_config())
    torch.set_default_dtype(torch.float32)
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * 0.02)
    model.requires_grad_(False)
    net = lora_ref.LoRANetworkRef(model, lora_dim=rank, target_class="WanTransformer3DModel", block_substr="blocks").to(dev)
    with torch.no_grad():
        for l in net.loras:
            l.lora_up.weight.normal_(0, 0.02)
    params = [p for l in net.loras for p in (l.lora_down.weight, l.lora_up.weight)]
    opt = torch.optim.AdamW(params, lr=1e-4, eps=1e-6)
    lat = torch.randn(batch, *WAN_LATENT, device=dev).bfloat16()
    noise = torch.randn_like(lat)
    t = torch.full((batch,), 500.0, device=dev)
    text = (torch.randn(batch, TEXT_LEN, 4096, device=dev) * 0.1).bfloat16()

    def step():
        opt.zero_grad(set_to_none=True)
        noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16)
        with net:
            loss = lora_ref.flow_loss(wan_ref.wan_predict(model, noisy, t, text), lat, noise)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    out = {"what": "eager PyTorch reference-style Wan2.1 step on the same GPU (oracle blocks + reference LoRA forward + torch "
                   "SDPA + torch AdamW), CUDA-event timed", "steps": steps, "warmup": warmup}
    for name, ck in (("checkpointing", True), ("no_checkpointing", False)):
        try:
            model.gradient_checkpointing = ck
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            out[name] = {"ms_per_step": e0.elapsed_time(e1) / steps, "loss": float(loss)}
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:100]}"}
            torch.cuda.empty_cache()
    out["peak_mem_gib"] = torch.cuda.max_memory_allocated() / 2 ** 30
    return out


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from ai_toolkit_b200 import LoRASpecialNetwork, cabi
    from ai_toolkit_b200.flux import FluxTransformer2DModel, flux_dev_config
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep

    ctx = cabi.Context.get(local)
    WAN = args.model == "wan"
    t_setup = time.time()
    BS, R = args.batch, args.rank
    if WAN:
        from ai_toolkit_b200 import wan_keys
        from ai_toolkit_b200.train_step import WanLoRATrainStep
        from ai_toolkit_b200.wan import WanTransformer3DModel, wan_1_3b_config

        cfg = wan_1_3b_config()
        if args.layers:
            cfg.num_layers = args.layers[0]
        model = WanTransformer3DModel(cfg, device=dev).init_synthetic_(seed=0)
        net = LoRASpecialNetwork(text_encoder=None, unet=model, lora_dim=R, alpha=R, train_unet=True, train_text_encoder=False,
                                 transformer_only=True, is_transformer=True, target_lin_modules=["WanTransformer3DModel"],
                                 base_model=wan_keys.WanLoRABaseModel())
    else:
        cfg = flux_dev_config()
        if args.layers:
            cfg.num_layers, cfg.num_single_layers = args.layers
        model = FluxTransformer2DModel(cfg, device=dev).init_synthetic_(seed=0)
        net = LoRASpecialNetwork(text_encoder=None, unet=model, lora_dim=R, alpha=R, train_unet=True,
                                 train_text_encoder=False, is_flux=True, transformer_only=True)
    net.force_to(dev, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    g = torch.Generator(device=dev).manual_seed(1234)
    with torch.no_grad():  # non-zero lora_up so that every gradient path carries signal (same on every rank)
        for m in net.get_all_modules():
            m.lora_up.weight.copy_(torch.randn(m.lora_up.weight.shape, generator=g, device=dev) * 0.02)
    net.mark_params_changed()
    opt = B200AdamW(net, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0, ema_decay=0.99,
                    grad_prescale=1.0 / world)
    hg = torch.Generator().manual_seed(1234 + rank)
    lat_shape = WAN_LATENT if WAN else LATENT
    if WAN:
        step = WanLoRATrainStep(model, net, opt, batch_size=BS, latent_shape=WAN_LATENT, text_len=TEXT_LEN,
                                use_cuda_graph=not args.no_graph)
    else:
        step = FluxLoRATrainStep(model, net, opt, batch_size=BS, latent_shape=LATENT, text_len=TEXT_LEN, guidance_scale=1.0,
                                 use_cuda_graph=not args.no_graph)
    # synthetic batch: host (pinned) copies for the e2e leg, seeded per rank
    host = {
        "latents": torch.randn(BS, *lat_shape, generator=hg).bfloat16().pin_memory(),
        "noise": torch.randn(BS, *lat_shape, generator=hg).bfloat16().pin_memory(),
        "timesteps": torch.linspace(1000, 1, 1000)[torch.randint(0, 999, (BS,), generator=hg)].float().pin_memory(),
        "text_embeds": (torch.randn(BS, TEXT_LEN, 4096, generator=hg) * 0.1).bfloat16().pin_memory(),
    }
    if not WAN:
        host["pooled_embeds"] = torch.randn(BS, 768, generator=hg).bfloat16().pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    step._load_dict(host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # launches of OUR kernels per step, counted on an eager step (graph replays do not pass through the C ABI)
    n0 = ctx.launch_count()
    step.run()
    torch.cuda.synchronize()
    launches_per_step = ctx.launch_count() - n0
    losses = []
    for _ in range(max(3, args.warmup)):  # includes the CUDA-graph capture
        losses.append(step.run())
    barrier()
    setup_s = time.time() - t_setup
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        loss_dev = step.run()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    # e2e: host batches in, loss out, every step
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(args.steps):
        last = step.hook_train_loop(host)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
    sampler.stop_flag = True
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # dominant kernel, timed live: the fused LoRA-Linear GEMM at the FLUX MLP-up shape
    peaks, peak_kind = measured_peaks()
    roof = None
    if rank == 0:
        M_, N_, K_ = (4608 if BS == 1 else 4096 * BS), 12288, 3072  # bs 4: the image stream of a double block, M = 16384
        if WAN:
            M_, N_, K_ = 13312 * BS, 8960, 1536  # ffn.net.0.proj of a Wan block
        x = (torch.randn(M_, K_, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N_, K_, device=dev) * 0.02).bfloat16()
        zc = (torch.randn(M_, 64, device=dev) * 0.1).bfloat16()
        bp = (torch.randn(N_, 64, device=dev) * 0.02).bfloat16()
        bias = torch.zeros(N_, device=dev, dtype=torch.bfloat16)
        y = torch.empty(M_, N_, device=dev, dtype=torch.bfloat16)
        pre = torch.empty(M_, N_, device=dev, dtype=torch.bfloat16)
        flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)

        def launch():  # ff.net.0.proj of a double block: base GEMM + LoRA segment + bias + GELU, pre-activation saved
            cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, act=cabi.ACT_GELU_TANH, aux_out=pre)

        for _ in range(3):
            launch()
        tt = []
        for _ in range(10):
            flush.zero_()  # L2 flush between timed launches (buffer > 126 MB L2)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            launch()
            b.record()
            torch.cuda.synchronize()
            tt.append(a.elapsed_time(b))
        kms = sum(tt) / len(tt)
        fl = 2.0 * M_ * N_ * K_ + 2.0 * M_ * R * N_  # base GEMM + rank-r up-projection riding in the same tile
        ach = fl / kms / 1e9
        prof = profile_numbers() if (BS == 1 and R == 16 and not WAN) else {"traffic": None, "source": None}
        roof = {"bound": "tensor",
                "kernel": f"gemm_bf16_kernel<2,256,6,0,0> fused LoRA-Linear + bias + GELU(+pre-activation) M={M_} N={N_} K={K_} r={R}",
                "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "peak_kind": f"{peak_kind} burst (kernel timed alone, L2 flushed)",
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch, parsed from the committed ncu --set full
                # summary named in `traffic_source` (null when no capture of this exact shape is committed)
                "traffic": prof.get("traffic"), "traffic_unit": "bytes/launch", "traffic_source": prof.get("source"),
                "algorithmic_bytes": 2.0 * (M_ * K_ + N_ * K_ + 2 * M_ * N_ + M_ * 64 + N_ * 64),
                "tensor_pipe_active_pct_ncu": prof.get("tensor_pipe_active_pct_ncu"), "us_per_launch": kms * 1e3}
    if rank != 0:
        return
    if WAN:
        f_step, f_lin, f_attn, f_lora = wan_flops(BS, R, layers=cfg.num_layers)
    else:
        f_step, f_lin, f_attn, f_lora = flux_flops(BS, R, n_double=cfg.num_layers, n_single=cfg.num_single_layers)
    value = world * 1e3 / ms
    out = {
        "metric": (f"train-steps/sec Wan2.1-T2V-1.3B LoRA r={R} bs={BS} 49x512x512" if WAN else
                   METRIC if (BS == 1 and R == 16) else f"train-steps/sec FLUX.1-dev LoRA r={R} bs={BS} 1024^2"), "value": value, "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": world * BS, "img_per_s": world * BS * 1e3 / ms,
                   "tokens": (13312 + 512 if WAN else 4608) * BS, "rank": R, "lora_modules": len(net.get_all_modules()), "lora_params": int(net.n_params),
                   "blocks": [cfg.num_layers] if WAN else [cfg.num_layers, cfg.num_single_layers], "parallelism": f"dp{world}", "ema": True,
                   "cuda_graph": not args.no_graph,
                   "l2": "per-step working set (23.8 GB weights + activations) >> 126 MB L2; no explicit flush"},
        "impl": "b200",
        "step_roofline": {"bound": "tensor", "f_step_tflop": f_step / 1e12, "achieved": f_step / (ms * 1e-3) / 1e12,
                          "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                          "frac": f_step / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"], "peak_kind": f"{peak_kind} sustained"},
        "roofline": roof,
        "e2e": {"value": world * 1e3 / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches_per_step) * args.steps,
        "launches_per_step": int(launches_per_step),
        "clocks": sampler.summary(),
        "loss_last": last["loss"] if last else None,
        "setup_s": setup_s,
    }
    if not args.skip_cpu_baseline and world == 1 and not WAN:  # rank 0 at N = 1 only (a bounded CPU sample of the same workload)
        sec, times, f_sample, desc, threads = cpu_reference_sample(steps=3, warmup=1, rank=R)
        v = 1.0 / (sec * f_step / f_sample)
        out["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": desc,
                               "step_seconds": times}
    if not args.skip_gpu_reference and world == 1 and not args.layers and WAN:
        del step, opt, net, model
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        ref = gpu_reference_leg_wan(dev, rank=R, batch=BS)
        for k in ("checkpointing", "no_checkpointing"):
            if "ms_per_step" in ref.get(k, {}):
                ref[k]["speedup_of_this_repo"] = ref[k]["ms_per_step"] / ms
        out["gpu_reference"] = ref
    elif not args.skip_gpu_reference and world == 1 and not args.layers:
        # free this arm's model / activations first: the eager reference needs its own 24 GB of weights + autograd state
        del step, opt, net, model
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        ref = gpu_reference_leg(dev, rank=R, batch=BS)
        for k in ("checkpointing", "no_checkpointing"):
            if "ms_per_step" in ref.get(k, {}):
                ref[k]["speedup_of_this_repo"] = ref[k]["ms_per_step"] / ms
        out["gpu_reference"] = ref
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# SDXL (BASELINE.json configs[1]): hybrid arm -- see ai_toolkit_b200/unet.py
# ---------------------------------------------------------------------------------------------------
def run_b200_sdxl(args):
    """configs[1]: SDXL-base LoRA r=8, 1024x1024, bs=2, one GPU.  The Transformer2DModel stacks run on this repo's kernels, the
    frozen ResNet / sampler body is eager PyTorch (cuDNN); `config.body` says so.  `gpu_reference` = the same step with the
    oracle's eager UNet (bf16, SDPA attention, adapters as forward hooks, torch AdamW) on the same GPU."""
    import torch

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--model sdxl is a single-GPU measurement (the UNet step has no replica plumbing)")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from ai_toolkit_b200 import LoRASpecialNetwork, cabi
    from ai_toolkit_b200 import unet as host_unet
    from ai_toolkit_b200.optimizer import B200AdamW

    ctx = cabi.Context.get(0)
    SD15 = args.model == "sd15"  # configs[0]: SD1.5 r=4, 512x512, bs=1 (the reference's CPU-runnable case, here on the GPU)
    BS = args.batch if (args.batch != 1 or SD15) else 2
    R = args.rank if args.rank != RANK else (4 if SD15 else 8)
    H = W = 64 if SD15 else 128
    cfg = host_unet.sd15_config() if SD15 else host_unet.sdxl_config()
    TD = cfg.cross_attention_dim
    model = host_unet.UNet2DConditionModel(cfg, device=dev).init_synthetic_(seed=0)
    net = LoRASpecialNetwork(text_encoder=None, unet=model, lora_dim=R, alpha=R, train_unet=True, train_text_encoder=False, is_sdxl=not SD15)
    net.force_to(dev, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    g = torch.Generator(device=dev).manual_seed(1234)
    with torch.no_grad():
        for m in net.get_all_modules():
            m.lora_up.weight.copy_(torch.randn(m.lora_up.weight.shape, generator=g, device=dev) * 0.02)
    net.mark_params_changed()
    opt = B200AdamW(net, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0)
    step = host_unet.UNetLoRATrainStep(model, net, opt, prediction_type="epsilon", use_cuda_graph=not args.no_graph)
    hg = torch.Generator().manual_seed(1234)
    host = {
        "latents": (torch.randn(BS, 4, H, W, generator=hg) * 0.18215 * 5).bfloat16().pin_memory(),
        "noise": torch.randn(BS, 4, H, W, generator=hg).bfloat16().pin_memory(),
        "timesteps": torch.randint(1, 999, (BS,), generator=hg).pin_memory(),
        "text_embeds": torch.randn(BS, 77, TD, generator=hg).bfloat16().pin_memory(),
    }
    if not SD15:
        host["pooled_embeds"] = torch.randn(BS, 1280, generator=hg).bfloat16().pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d = {k: v.to(dev) for k, v in host.items()}
    step.load_batch(host["latents"], host["noise"], host["timesteps"], host["text_embeds"], host.get("pooled_embeds"))

    def run_dev():  # the resident batch (static device buffers), as the FLUX arm does
        return step.run()

    n0 = ctx.launch_count()
    run_dev()
    torch.cuda.synchronize()
    launches_per_step = ctx.launch_count() - n0
    for _ in range(max(3, args.warmup)):
        run_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        run_dev()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = step.hook_train_loop(host)  # pinned host batch -> static device buffers -> step -> loss to the host
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.stop_flag = True
    peaks, peak_kind = measured_peaks()
    f_step = (host_unet.SD15_STEP_FLOPS_PER_SAMPLE if SD15 else host_unet.SDXL_STEP_FLOPS_PER_SAMPLE) * BS
    out = {
        "metric": (f"train-steps/sec SD1.5 LoRA r={R} bs={BS} 512^2" if SD15 else f"train-steps/sec SDXL-base LoRA r={R} bs={BS} 1024^2"),
        "value": 1e3 / ms, "unit": "steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("BASELINE.json configs[0]: SD1.5 LoRA r=4 bs=1 512x512 (latents 1x4x64x64, text 1x77x768), bf16 on the GPU"
                                if SD15 else
                                "BASELINE.json configs[1]: SDXL-base LoRA r=8 bs=2 1024x1024 (latents 2x4x128x128, text 2x77x2048)"),
                   "global_batch": BS, "img_per_s": BS * 1e3 / ms, "rank": R, "lora_modules": len(net.get_all_modules()),
                   "lora_params": int(net.n_params), "parallelism": "dp1", "cuda_graph": not args.no_graph,
                   "body": "HYBRID: Transformer2DModel stacks (adapter-bearing) on this repo's kernels; frozen ResnetBlock2D / "
                           "Down/Upsample2D / conv_in/out / time embeddings = eager PyTorch (cuDNN / cuBLAS) under autograd",
                   "l2": "per-step working set (1.7 GB (SD1.5) / 5.1 GB (SDXL) of weights + activations) >> 126 MB L2; no explicit flush"},
        "impl": "b200",
        "step_roofline": {"bound": "tensor", "f_step_tflop": f_step / 1e12, "achieved": f_step / (ms * 1e-3) / 1e12,
                          "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                          "frac": f_step / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"], "peak_kind": f"{peak_kind} sustained",
                          "f_step_source": "FlopCounterMode over oracle/unet_ref.py fwd+bwd (SURVEY.md section 8d)"},
        "roofline": None,
        "e2e": {"value": 1e3 / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches_per_step) * args.steps, "launches_per_step": int(launches_per_step),
        "clocks": sampler.summary(), "loss_last": last["loss"] if last else None,
    }
    if not args.skip_gpu_reference:
        del step, opt, net, model
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        out["gpu_reference"] = gpu_reference_leg_sdxl(dev, d, R, ms, sd15=SD15)
    print(json.dumps(out), flush=True)


def gpu_reference_leg_sdxl(dev, d, R, ms_ours, steps=3, warmup=2, sd15=False):
    import torch
    import torch.nn.functional as F

    from oracle import unet_ref

    cfg = unet_ref.sd15_config() if sd15 else unet_ref.sdxl_config()
    torch.manual_seed(0)
    with torch.device(dev):
        om = unet_ref.UNet2DConditionModel(cfg).to(torch.bfloat16)
    with torch.no_grad():
        for name, p in om.named_parameters():
            p.copy_(torch.randn(p.shape, device=dev) * 0.02 if not (("norm" in name) and name.endswith(".weight") and p.dim() == 1)
                    else 1.0 + 0.1 * torch.randn(p.shape, device=dev))
    om.requires_grad_(False)
    params = []
    for name, mod in om.named_modules():
        if ".attentions." not in name:
            continue
        is_lin = isinstance(mod, torch.nn.Linear)
        is_c1 = isinstance(mod, torch.nn.Conv2d) and mod.kernel_size == (1, 1)
        if not (is_lin or is_c1):
            continue
        cin = mod.in_features if is_lin else mod.in_channels
        cout = mod.out_features if is_lin else mod.out_channels
        A = torch.nn.Parameter(torch.randn(R, cin, device=dev) * 0.02)
        Bw = torch.nn.Parameter(torch.randn(cout, R, device=dev) * 0.02)
        params += [A, Bw]

        def hook(m, inp, out, A=A, Bw=Bw):  # the reference's LoRAModule.forward under bf16 autocast: org + up(down(x)) * scale
            x = inp[0]
            if x.dim() == 4:  # 1x1-conv projection (SD1.5)
                return out + F.conv2d(F.conv2d(x, A.to(x.dtype)[:, :, None, None]), Bw.to(x.dtype)[:, :, None, None])
            return out + F.linear(F.linear(x, A.to(x.dtype)), Bw.to(x.dtype))

        mod.register_forward_hook(hook)
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-2)
    B, _, H, W = d["latents"].shape
    tid = torch.tensor([[H * 8, W * 8, 0, 0, H * 8, W * 8]] * B, device=dev, dtype=torch.float32)
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, device=dev) ** 2
    ac = torch.cumprod(1 - betas, 0)

    def one():
        opt.zero_grad(set_to_none=True)
        a = ac[d["timesteps"]]
        noisy = (a.sqrt()[:, None, None, None] * d["latents"].float() + (1 - a).sqrt()[:, None, None, None] * d["noise"].float()).bfloat16()
        pred = om(noisy, d["timesteps"].float(), d["text_embeds"],
                  added_cond_kwargs=None if sd15 else {"text_embeds": d["pooled_embeds"], "time_ids": tid})[0]
        loss = ((pred.float() - d["noise"].float()) ** 2).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    try:
        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"what": "oracle/unet_ref.py UNet, bf16 eager + SDPA, adapters as forward hooks, torch AdamW, no checkpointing",
                "ms_per_step": ms, "steps": steps, "loss": float(loss.detach()), "speedup_of_this_repo": ms / ms_ours,
                "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    except Exception as e:  # noqa: BLE001 -- a reported side measurement must not take the main line down
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-gpu-reference", action="store_true",
                    help="do not time the eager reference-style PyTorch step on the same GPU after the main measurement")
    ap.add_argument("--model", default="flux", choices=["flux", "wan", "sdxl", "sd15"],
                    help="flux = BASELINE.json configs[2] (the headline metric); wan = configs[3] (Wan2.1-T2V-1.3B, 49x512x512); "
                         "sdxl = configs[1], sd15 = configs[0] (hybrid: engine blocks + eager frozen body, one GPU)")
    ap.add_argument("--batch", type=int, default=1, help="samples per GPU (BASELINE.json configs[4] uses 4)")
    ap.add_argument("--rank", type=int, default=RANK, help="LoRA rank (configs[4] sweeps 4, 8, 16, 32, 64)")
    ap.add_argument("--layers", type=int, nargs=2, default=None, help="debug: override (double, single) block counts")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.model in ("sdxl", "sd15"):
        run_b200_sdxl(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
