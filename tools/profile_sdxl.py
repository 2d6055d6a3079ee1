"""Per-kernel GPU time of one SDXL hybrid step (torch.profiler / CUPTI; eager launches, so the kernel times are the step's GPU
work without the host-launch gaps).  Output: top kernels by total time.  Not a bench number (profiler attached)."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_toolkit_b200 import LoRASpecialNetwork  # noqa: E402
from ai_toolkit_b200 import unet as host_unet  # noqa: E402
from ai_toolkit_b200.optimizer import B200AdamW  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = host_unet.sdxl_config()
model = host_unet.UNet2DConditionModel(cfg, device=dev).init_synthetic_(seed=0)
if os.environ.get("CHANNELS_LAST") == "1":
    model = model.to(memory_format=torch.channels_last)
net = LoRASpecialNetwork(text_encoder=None, unet=model, lora_dim=8, alpha=8, train_unet=True, train_text_encoder=False, is_sdxl=True)
net.force_to(dev, torch.float32)
net._update_torch_multiplier()
net.apply_to(None, model, False, True)
g = torch.Generator(device=dev).manual_seed(1)
with torch.no_grad():
    for m in net.get_all_modules():
        m.lora_up.weight.copy_(torch.randn(m.lora_up.weight.shape, generator=g, device=dev) * 0.02)
net.mark_params_changed()
opt = B200AdamW(net, lr=1e-4, max_grad_norm=1.0)
step = host_unet.UNetLoRATrainStep(model, net, opt)
B = 2
hg = torch.Generator().manual_seed(2)
step.load_batch(torch.randn(B, 4, 128, 128, generator=hg).bfloat16(), torch.randn(B, 4, 128, 128, generator=hg).bfloat16(),
                torch.randint(1, 999, (B,), generator=hg), torch.randn(B, 77, 2048, generator=hg).bfloat16(),
                torch.randn(B, 1280, generator=hg).bfloat16())
for _ in range(3):
    step.run()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step.run()
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:110]
        tot[k][0] += 1
        tot[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
total = sum(v[1] for v in tot.values())
print(f"total GPU kernel time {total / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} kernels")
ours = sum(v[1] for k, v in tot.items() if "b200::" in k)
print(f"this repo's kernels: {ours / 1e3:.2f} ms ({100 * ours / total:.1f} %)")
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {k} | {n} | {t / 1e3:.2f} | {100 * t / total:.1f}% | {t / n:.1f} |")
