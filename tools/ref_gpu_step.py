"""R-GPU row of BASELINE.md section 4: the reference-style EAGER PyTorch step on one B200 (informational; the bench's
`--impl reference` arm is the CPU one).  Oracle FLUX blocks (diffusers names) + the oracle restatement of the
reference's LoRA forward (fp32 side branch, network_mixins.py:304-342) + torch SDPA + torch.optim.AdamW(eps=1e-6) +
clip_grad_norm_(1.0), bf16 base / fp32 LoRA, synthetic weights and batch, with and without gradient checkpointing.
usage: python tools/ref_gpu_step.py [ckpt|nockpt] [steps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import flux_ref, lora_ref  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "ckpt"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    model = flux_ref.FluxTransformer2DModel(flux_ref.flux_dev_config())
torch.set_default_dtype(torch.float32)
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    for p in model.parameters():
        p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * 0.02)
model.requires_grad_(False)
model.gradient_checkpointing = mode == "ckpt"
net = lora_ref.LoRANetworkRef(model, lora_dim=16).to(dev)
with torch.no_grad():
    for l in net.loras:
        l.lora_up.weight.normal_(0, 0.02)
params = [p for l in net.loras for p in (l.lora_down.weight, l.lora_up.weight)]
opt = torch.optim.AdamW(params, lr=1e-4, eps=1e-6)
lat = torch.randn(1, 16, 128, 128, device=dev).bfloat16()
noise = torch.randn_like(lat)
t = torch.tensor([500.0], device=dev)
text = (torch.randn(1, 512, 4096, device=dev) * 0.1).bfloat16()
pooled = torch.randn(1, 768, device=dev).bfloat16()


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


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"R-GPU eager reference-style step ({'gradient checkpointing' if mode == 'ckpt' else 'no checkpointing'}): "
      f"{ms:.1f} ms/step = {1e3 / ms:.3f} steps/s; loss {float(loss):.4f}; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
