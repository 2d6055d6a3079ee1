"""TEST INFRASTRUCTURE ONLY — generate tests/golden/lora_tiny.pt by running the UNMODIFIED reference
(`/root/reference`, ostris/ai-toolkit @ 27a03a9: toolkit.lora_special.LoRASpecialNetwork + network_mixins forward +
torch.optim.AdamW(eps=1e-6) + clip_grad_norm_) on the oracle's tiny FLUX-structured model, CPU fp32, fixed seeds.
Run here (the reference tree does not exist on the GPU box):   python oracle/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flux_ref, lora_ref, ref_import  # noqa: E402


def build(seed=0):
    torch.manual_seed(seed)
    cfg = flux_ref.FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=64,
                              pooled_projection_dim=32)
    model = flux_ref.init_synthetic_(flux_ref.FluxTransformer2DModel(cfg), seed=seed, std=0.05).requires_grad_(False)
    g = torch.Generator().manual_seed(seed + 1)
    batch = dict(latents=torch.randn(2, 16, 8, 8, generator=g), noise=torch.randn(2, 16, 8, 8, generator=g),
                 timesteps=torch.tensor([500.0, 250.0]), text=torch.randn(2, 8, 64, generator=g) * 0.5,
                 pooled=torch.randn(2, 32, generator=g))
    return cfg, model, batch


def main():
    RefNet, _ = ref_import.reference_lora()
    cfg, model, batch = build()
    torch.manual_seed(123)
    net = RefNet(text_encoder=None, unet=model, lora_dim=4, alpha=4, train_unet=True, train_text_encoder=False, is_flux=True,
                 network_type="lora", transformer_only=True)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for lora in net.unet_loras:
            lora.lora_up.weight.copy_(torch.randn(lora.lora_up.weight.shape, generator=g) * 0.05)
    init_sd = {k: v.clone() for k, v in net.state_dict().items()}
    params = net.prepare_optimizer_params(1e-3, 1e-3, 1e-3)
    opt = torch.optim.AdamW(params, lr=1e-3, eps=1e-6)
    losses = []
    grads0 = None
    pred0 = None
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        noisy = lora_ref.add_noise_flowmatch(batch["latents"], batch["noise"], batch["timesteps"])
        with net:
            pred = lora_ref.flux_predict(model, noisy, batch["timesteps"], batch["text"], batch["pooled"], 1.0,
                                         flux_ref.pack_latents, flux_ref.unpack_latents, flux_ref.make_img_ids)
            loss = lora_ref.flow_loss(pred, batch["latents"], batch["noise"])
            loss.backward()
        if step == 0:
            pred0 = pred.detach().clone()
            grads0 = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
        torch.nn.utils.clip_grad_norm_([p for g_ in params for p in g_["params"]], 1.0)
        opt.step()
        losses.append(float(loss))
    out = dict(batch=batch, init_state_dict=init_sd, pred0=pred0, grads0=grads0, losses=losses,
               final_state_dict={k: v.clone() for k, v in net.state_dict().items()},
               saved_keys=list(net.get_state_dict(dtype=torch.float16).keys()),
               lora_names=[l.lora_name for l in net.unet_loras], reference_commit="27a03a9")
    path = os.path.join(ROOT, "tests", "golden", "lora_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; losses", losses)


if __name__ == "__main__":
    main()
