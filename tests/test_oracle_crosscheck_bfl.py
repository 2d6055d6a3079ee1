"""Independent cross-check of the oracle's FLUX DiT (oracle/flux_ref.py, a restatement of diffusers'
`FluxTransformer2DModel`, which the reference calls but does not vendor — DESIGN.md "parity unpinned at the diffusers
boundary").  An unrelated BFL-lineage implementation of the same published model happens to be installed with
torchtitan (`torchtitan/experiments/flux/model`, third-party, not authoritative).  With the public BFL -> diffusers weight
conversion (fused qkv split into to_q/to_k/to_v, `linear1` into to_q/to_k/to_v/proj_mlp, `*_mod.lin` -> `norm*.linear`,
final-layer shift/scale swapped) the two must compute the same function; they do, in fp32 on CPU, forward and backward.
This does not pin the oracle to the reference (that needs diffusers), it removes the possibility that the restatement
is a private misreading of the architecture.  Skipped where torchtitan is absent."""
import pytest
import torch

from oracle import flux_ref

tt = pytest.importorskip("torchtitan.experiments.flux.model.model")
tt_args = pytest.importorskip("torchtitan.experiments.flux.model.args")


def _convert(bfl_sd, n_double, n_single, d, mlp):
    """BFL / torchtitan parameter names -> diffusers names (the published conversion; no guidance embedder here)."""
    sd = {}
    sd["x_embedder.weight"], sd["x_embedder.bias"] = bfl_sd["img_in.weight"], bfl_sd["img_in.bias"]
    sd["context_embedder.weight"], sd["context_embedder.bias"] = bfl_sd["txt_in.weight"], bfl_sd["txt_in.bias"]
    for a, b in (("time_in", "timestep_embedder"), ("vector_in", "text_embedder")):
        for i, l in ((1, "in_layer"), (2, "out_layer")):
            for wb in ("weight", "bias"):
                sd[f"time_text_embed.{b}.linear_{i}.{wb}"] = bfl_sd[f"{a}.{l}.{wb}"]
    for i in range(n_double):
        s, t = f"double_blocks.{i}.", f"transformer_blocks.{i}."
        for wb in ("weight", "bias"):
            sd[t + f"norm1.linear.{wb}"] = bfl_sd[s + f"img_mod.lin.{wb}"]
            sd[t + f"norm1_context.linear.{wb}"] = bfl_sd[s + f"txt_mod.lin.{wb}"]
            q, k, v = bfl_sd[s + f"img_attn.qkv.{wb}"].chunk(3, 0)
            sd[t + f"attn.to_q.{wb}"], sd[t + f"attn.to_k.{wb}"], sd[t + f"attn.to_v.{wb}"] = q, k, v
            q, k, v = bfl_sd[s + f"txt_attn.qkv.{wb}"].chunk(3, 0)
            sd[t + f"attn.add_q_proj.{wb}"], sd[t + f"attn.add_k_proj.{wb}"], sd[t + f"attn.add_v_proj.{wb}"] = q, k, v
            sd[t + f"attn.to_out.0.{wb}"] = bfl_sd[s + f"img_attn.proj.{wb}"]
            sd[t + f"attn.to_add_out.{wb}"] = bfl_sd[s + f"txt_attn.proj.{wb}"]
            sd[t + f"ff.net.0.proj.{wb}"] = bfl_sd[s + f"img_mlp.0.{wb}"]
            sd[t + f"ff.net.2.{wb}"] = bfl_sd[s + f"img_mlp.2.{wb}"]
            sd[t + f"ff_context.net.0.proj.{wb}"] = bfl_sd[s + f"txt_mlp.0.{wb}"]
            sd[t + f"ff_context.net.2.{wb}"] = bfl_sd[s + f"txt_mlp.2.{wb}"]
        sd[t + "attn.norm_q.weight"] = bfl_sd[s + "img_attn.norm.query_norm.weight"]
        sd[t + "attn.norm_k.weight"] = bfl_sd[s + "img_attn.norm.key_norm.weight"]
        sd[t + "attn.norm_added_q.weight"] = bfl_sd[s + "txt_attn.norm.query_norm.weight"]
        sd[t + "attn.norm_added_k.weight"] = bfl_sd[s + "txt_attn.norm.key_norm.weight"]
    for i in range(n_single):
        s, t = f"single_blocks.{i}.", f"single_transformer_blocks.{i}."
        for wb in ("weight", "bias"):
            sd[t + f"norm.linear.{wb}"] = bfl_sd[s + f"modulation.lin.{wb}"]
            q, k, v, m = bfl_sd[s + f"linear1.{wb}"].split([d, d, d, mlp], 0)
            sd[t + f"attn.to_q.{wb}"], sd[t + f"attn.to_k.{wb}"], sd[t + f"attn.to_v.{wb}"] = q, k, v
            sd[t + f"proj_mlp.{wb}"] = m
            sd[t + f"proj_out.{wb}"] = bfl_sd[s + f"linear2.{wb}"]
        sd[t + "attn.norm_q.weight"] = bfl_sd[s + "norm.query_norm.weight"]
        sd[t + "attn.norm_k.weight"] = bfl_sd[s + "norm.key_norm.weight"]
    for wb in ("weight", "bias"):
        sd[f"proj_out.{wb}"] = bfl_sd[f"final_layer.linear.{wb}"]
        shift, scale = bfl_sd[f"final_layer.adaLN_modulation.1.{wb}"].chunk(2, 0)
        sd[f"norm_out.linear.{wb}"] = torch.cat([scale, shift], 0)  # diffusers' AdaLayerNormContinuous: scale first
    return sd


def test_oracle_flux_equals_bfl_lineage_implementation():
    torch.manual_seed(0)
    heads, d, n_double, n_single, ctx, vec = 2, 256, 2, 2, 48, 32
    args = tt_args.FluxModelArgs(in_channels=64, out_channels=64, vec_in_dim=vec, context_in_dim=ctx, hidden_size=d,
                                 mlp_ratio=4.0, num_heads=heads, depth=n_double, depth_single_blocks=n_single,
                                 axes_dim=(16, 56, 56), qkv_bias=True)
    bfl = tt.FluxModel(args).float()
    with torch.no_grad():  # random everywhere (the library's own init zeroes the modulations and the output layer)
        for n, p in bfl.named_parameters():
            if n.endswith("norm.weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            else:
                p.normal_(0, 0.05)
    cfg = flux_ref.FluxConfig(in_channels=64, num_layers=n_double, num_single_layers=n_single, attention_head_dim=128,
                              num_attention_heads=heads, joint_attention_dim=ctx, pooled_projection_dim=vec,
                              guidance_embeds=False)
    ora = flux_ref.FluxTransformer2DModel(cfg).float()
    conv = _convert(dict(bfl.state_dict()), n_double, n_single, d, 4 * d)
    own = ora.state_dict()
    assert set(conv) == set(own), (sorted(set(own) - set(conv))[:5], sorted(set(conv) - set(own))[:5])
    ora.load_state_dict(conv)

    B, hl, wl, Lt = 2, 8, 12, 10
    img = torch.randn(B, (hl // 2) * (wl // 2), 64, requires_grad=True)
    txt = torch.randn(B, Lt, ctx)
    y = torch.randn(B, vec)
    t = torch.tensor([0.25, 0.9])
    img_ids = flux_ref.make_img_ids(hl, wl)
    txt_ids = torch.zeros(Lt, 3)
    want = bfl(img=img, img_ids=img_ids[None].expand(B, -1, -1), txt=txt, txt_ids=txt_ids[None].expand(B, -1, -1),
               timesteps=t, y=y)
    g_want, = torch.autograd.grad(want.square().sum(), img)
    img2 = img.detach().clone().requires_grad_(True)
    got = ora(img2, t, txt, y, txt_ids, img_ids, guidance=None)
    g_got, = torch.autograd.grad(got.square().sum(), img2)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()  # noqa: E731
    assert rel(got, want) < 2e-5, rel(got, want)
    assert rel(g_got, g_want) < 2e-4, rel(g_got, g_want)
