"""Run the full FLUX.1-dev LoRA step eagerly (no CUDA graph) with cudaProfilerStart/Stop around ONE step, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` (launch list) — see tools/gpu_profile_trip.sh."""
import sys

import torch

sys.path.insert(0, ".")
from ai_toolkit_b200 import LoRASpecialNetwork  # noqa: E402
from ai_toolkit_b200.flux import FluxTransformer2DModel, flux_dev_config  # noqa: E402
from ai_toolkit_b200.optimizer import B200AdamW  # noqa: E402
from ai_toolkit_b200.train_step import FluxLoRATrainStep  # noqa: E402

dev = torch.device("cuda:0")
cfg = flux_dev_config()
if len(sys.argv) > 2:
    cfg.num_layers, cfg.num_single_layers = int(sys.argv[1]), int(sys.argv[2])
model = FluxTransformer2DModel(cfg, device=dev).init_synthetic_(0)
net = LoRASpecialNetwork(None, model, lora_dim=16, alpha=16, train_text_encoder=False, is_flux=True, transformer_only=True)
net.force_to(dev, torch.float32)
net._update_torch_multiplier()
net.apply_to(None, model, False, True)
with torch.no_grad():
    for m in net.get_all_modules():
        m.lora_up.weight.normal_(0, 0.02)
net.mark_params_changed()
opt = B200AdamW(net, lr=1e-4, ema_decay=0.99)
step = FluxLoRATrainStep(model, net, opt, batch_size=1, use_cuda_graph=False)
step.latents.normal_()
step.noise.normal_()
step.timesteps.fill_(500.0)
step.text.normal_(0, 0.1)
step.pooled.normal_()
step.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(step.loss_ws[1]))
