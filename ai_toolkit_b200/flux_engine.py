"""Hand-scheduled forward / backward of the frozen FLUX DiT with LoRA on every block Linear.

This replaces autograd over `diffusers.FluxTransformer2DModel` + 494 `ToolkitModuleMixin.forward` hooks
 by an explicit launch schedule of the C-ABI kernels:

  * every Linear = `linear.linear_fwd / linear_bwd` (one fused tcgen05 GEMM + one rank-side GEMM forward;
    dgrad with the frozen weight consumed MN-major, dA / dB by token-contraction GEMMs into the flat grad buffer)
  * AdaLN-Zero modulation, QK-RMSNorm + RoPE + head re-layout, gate/residual, GELU: fused row kernels or GEMM
    epilogues (`elementwise.cu`, `gemm_tcgen05.cu`)
  * joint attention: `attention.fwd / bwd`
  * the AdaLN projections of the [B, D] conditioning vector: weight-streaming GEMV kernels (`gemv.cu`)

Activations needed by the backward are kept (no recompute): ~31 GB for FLUX.1-dev at 1024^2, bs 1, which
is what 180 GB of HBM is for; the reference's default gradient checkpointing re-runs the forward instead.
The whole schedule is static (shapes, pointers), so one training step can be captured in a CUDA graph
(`train_step.FluxLoRATrainStep`).

Model arithmetic follows diffusers' FLUX blocks as restated in oracle/flux_ref.py (in-tree anchors:
extensions_built_in/diffusion_models/chroma/src/layers.py:471-681, math.py:13-51).
"""
from __future__ import annotations

import torch

from . import attention, cabi, ops
from .cabi import ACT_GELU_TANH, gemm_bf16
from .linear import group_bwd, group_fwd, linear_bwd, linear_fwd, live_lora, lora_coeff


def _empty(shape, like, dtype=torch.bfloat16):
    return torch.empty(shape, device=like.device, dtype=dtype)


class FluxEngine:
    def __init__(self, model):
        self.model = model
        cfg = model.cfg
        self.D = cfg.inner_dim
        self.H = cfg.num_attention_heads
        assert cfg.attention_head_dim == 128, "the attention / rope kernels are specialised for head_dim 128"
        self._rope_cache = {}
        self.saved = None

    # ------------------------------------------------------------------------------------------
    # small helpers
    # ------------------------------------------------------------------------------------------
    def rope_tables(self, txt_ids, img_ids):
        """FluxPosEmbed: float64 angles per axis, cos/sin repeat-interleaved to head_dim, fp32 tables [L, 128].
        Constant for a given resolution, so it is built once and cached (plumbing, not hot path)."""
        key = (txt_ids.data_ptr(), img_ids.data_ptr(), txt_ids.shape[0], img_ids.shape[0], str(img_ids.device))
        hit = self._rope_cache.get(key)
        if hit is not None:
            return hit
        ids = torch.cat((txt_ids, img_ids), dim=0).to(torch.float64)
        cos, sin = [], []
        for i, d in enumerate(self.model.cfg.axes_dims_rope):
            freqs = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device) / d))
            ang = torch.outer(ids[:, i], freqs)
            cos.append(ang.cos().repeat_interleave(2, dim=1).float())
            sin.append(ang.sin().repeat_interleave(2, dim=1).float())
        out = (torch.cat(cos, -1).contiguous(), torch.cat(sin, -1).contiguous())
        self._rope_cache[key] = out
        return out

    # Linears outside the transformer blocks: computed straight from their frozen weights by this engine
    _PRELUDE = ("x_embedder", "context_embedder", "time_text_embed.timestep_embedder.linear_1",
                "time_text_embed.timestep_embedder.linear_2", "time_text_embed.text_embedder.linear_1",
                "time_text_embed.text_embedder.linear_2", "time_text_embed.guidance_embedder.linear_1",
                "time_text_embed.guidance_embedder.linear_2", "norm_out.linear", "proj_out")

    def active_network(self):
        """The LoRASpecialNetwork applied to this model, or None.  Found through the model (`apply_to` leaves a weak
        reference) or, failing that, through ANY adapted Linear -- never through one particular layer, which
        only_if_contains / ignore_if_contains configs may leave unadapted."""
        ref = getattr(self.model, "_b200_network", None)
        net = ref() if ref is not None else None
        if net is None:
            for mod in self.model.modules():
                r = getattr(mod, "_b200_lora", None)
                lora = r() if r is not None else None
                if lora is not None and lora.network_ref is not None and lora.network_ref() is not None:
                    net = lora.network_ref()
                    break
        if net is None:
            return None
        if getattr(self, "_checked_for", None) is not net:
            # transformer_only=False (the class default; the trainer's NetworkConfig default is True) also wraps the
            # embedders / norm_out / proj_out: this engine has no adapter path (nor the conditioning-vector backward they
            # would need) for them.  Refuse instead of silently training without them.
            for name in self._PRELUDE:
                mod = self.model
                try:
                    for part in name.split("."):
                        mod = getattr(mod, part)
                except AttributeError:
                    continue
                if getattr(mod, "_b200_lora", None) is not None and mod._b200_lora() is not None:
                    raise NotImplementedError(
                        f"LoRA on `{name}` (transformer_only=False): the fused FLUX engine adapts the Linears under "
                        "transformer_blocks / single_transformer_blocks only; build the network with transformer_only=True "
                        "(the reference's NetworkConfig default, toolkit/config_modules.py:202)")
            self._checked_for = net
        return net

    def _register_groups(self, net):
        """q/k/v projections share their input: fuse each triple into one GEMM (weights re-pointed at one matrix)."""
        m = self.model
        groups = []
        for blk in m.transformer_blocks:
            a = blk.attn
            groups.append([live_lora(l) for l in (a.to_q, a.to_k, a.to_v)])
            groups.append([live_lora(l) for l in (a.add_q_proj, a.add_k_proj, a.add_v_proj)])
        for blk in m.single_transformer_blocks:
            a = blk.attn
            quad = [live_lora(l) for l in (blk.proj_mlp, a.to_q, a.to_k, a.to_v)]
            if all(l is not None for l in quad) and quad[0].lora_dim % 8 == 0 and 4 * quad[0].lora_dim <= 64:
                groups.append(quad)  # [proj_mlp; q; k; v] share the AdaLN output: one GEMM, N = 4D + 3D
            else:
                groups.append(quad[1:])
        net.register_fused_groups(groups)
        self._groups_for = net

    @staticmethod
    def _qkv_fwd(lins, n, qkv):
        """Three projections of the same input -> column blocks of `qkv`; fused when the adapters form a group."""
        loras = [live_lora(l) for l in lins]
        grp = loras[0].network_ref().fused_group(loras) if loras[0] is not None else None
        if grp is not None:
            return ("g", group_fwd(grp, lins, n, qkv))
        D = lins[0].out_features
        return ("s", [linear_fwd(lin, n, qkv[:, j * D:(j + 1) * D], lora=lo) for j, (lin, lo) in enumerate(zip(lins, loras))])

    @staticmethod
    def _qkv_bwd(lins, dqkv, n, saved, dn, accumulate):
        loras = [live_lora(l) for l in lins]
        kind, z = saved
        if kind == "g":
            grp = loras[0].network_ref().fused_group(loras)
            group_bwd(grp, lins, dqkv, n, z, dn, **({"res": dn} if accumulate else {}))
            return
        D = lins[0].out_features
        for j, (lin, lo) in enumerate(zip(lins, loras)):
            linear_bwd(lin, dqkv[:, j * D:(j + 1) * D], n, z[j], dn, lora=lo, **({"res": dn} if (accumulate or j > 0) else {}))

    @staticmethod
    def _lin(lin, x, out, **epi):
        return linear_fwd(lin, x, out, lora=live_lora(lin), **epi)

    @staticmethod
    def _mod_fwd(lin, temb_silu):
        """AdaLN projection of the conditioning vector: -> (mod [B, k*D] bf16, z fp32 | None, alpha)."""
        lora = live_lora(lin)
        if lora is None:
            y, _ = ops.lora_gemv_fwd(temb_silu, lin.weight, lin.bias)
            return y, None, (0.0, None)
        # per-sample multipliers (SDTrainer.py:1558 -> network_mixins.py:311-322): one row of the conditioning vector per
        # sample, so the per-sample coefficient is a per-row coefficient of the GEMV (alpha, row_c) saved for the backward
        alpha, row_alpha, rps = lora_coeff(lora, temb_silu.shape[0])
        row_c = None
        if row_alpha is not None:
            row_c = row_alpha.repeat_interleave(rps).contiguous() if rps > 1 else row_alpha
        y, z = ops.lora_gemv_fwd(temb_silu, lin.weight, lin.bias, lora.down_weight_2d(), lora.up_weight_2d(), alpha,
                                 row_c=row_c)
        return y, z, (alpha, row_c)

    @staticmethod
    def _mod_bwd(lin, dmod, temb_silu, z, alpha):
        lora = live_lora(lin)
        if lora is None:
            return
        alpha, row_c = alpha
        ops.lora_gemv_bwd(dmod, temb_silu, z, lora.down_weight_2d(), lora.up_weight_2d(), alpha,
                          lora.lora_down.weight.grad.view(lora.lora_dim, lora.in_dim),
                          lora.lora_up.weight.grad.view(lora.out_dim, lora.lora_dim), row_c=row_c)

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def forward(self, packed, t01, enc, pooled, guidance, txt_ids, img_ids, save=True, t_div=1.0):
        """packed [B, Li, 64] bf16; t01 [B] fp32 (= timestep / 1000, or the raw timestep with t_div=1000); enc [B, Lt, 4096] bf16; pooled [B, 768] bf16;
        guidance [B] fp32 or None.  Returns pred [B*Li, 64] bf16.  With save=True the activations for `backward`
        are kept in self.saved."""
        m = self.model
        D, H = self.D, self.H
        B, Li, Cin = packed.shape
        Lt = enc.shape[1]
        L = Lt + Li
        dev = packed.device
        if dev.type != "cuda":
            raise cabi.B200Error("FluxEngine needs a B200; there is no CPU / eager fallback")
        net = self.active_network()
        if net is not None and net.is_active and not net.is_merged_in and len(net.get_all_modules()):
            if getattr(self, "_groups_for", None) is not net:
                self._register_groups(net)
            if any(mod.has_dropout() or (mod.module_dropout and mod.training) for mod in net.get_all_modules()):
                raise NotImplementedError("dropout / rank_dropout / module_dropout with the fused FLUX engine: use the "
                                          "per-module path (LoRAModule.forward) or set them to None (the reference default)")
            net.refresh_packs()
            net.ensure_grad_views()
        cos, sin = self.rope_tables(txt_ids, img_ids)
        sv = {"B": B, "Li": Li, "Lt": Lt, "cos": cos, "sin": sin, "double": [], "single": []} if save else None

        # --- conditioning vector (frozen prelude; M = B rows -> weight-streaming GEMV kernels)
        tte = m.time_text_embed

        def mlp2(emb, x):
            h, _ = ops.lora_gemv_fwd(x, emb.linear_1.weight, emb.linear_1.bias)
            h = ops.silu(h)
            y, _ = ops.lora_gemv_fwd(h, emb.linear_2.weight, emb.linear_2.bias)
            return y

        t_emb = mlp2(tte.timestep_embedder, ops.timestep_embed(t01, 256, div=t_div, mult=1000.0))
        p_emb = mlp2(tte.text_embedder, pooled.contiguous())
        if m.cfg.guidance_embeds:
            g_emb = mlp2(tte.guidance_embedder, ops.timestep_embed(guidance, 256, div=1.0, mult=1000.0))
            temb = ops.add_bf16(t_emb, g_emb, p_emb)
        else:
            temb = ops.add_bf16(t_emb, p_emb)
        temb_silu = ops.silu(temb)

        img = _empty((B * Li, D), packed)
        gemm_bf16(packed.reshape(B * Li, Cin), m.x_embedder.weight, img, bias=m.x_embedder.bias)
        txt = _empty((B * Lt, D), packed)
        gemm_bf16(enc.reshape(B * Lt, enc.shape[2]), m.context_embedder.weight, txt, bias=m.context_embedder.bias)

        for blk in m.transformer_blocks:
            img, txt, s = self._double_fwd(blk, img, txt, temb_silu, B, Li, Lt, cos, sin, save)
            if save:
                sv["double"].append(s)

        x = _empty((B * L, D), packed)
        xv = x.view(B, L, D)
        xv[:, :Lt].copy_(txt.view(B, Lt, D))
        xv[:, Lt:].copy_(img.view(B, Li, D))
        for blk in m.single_transformer_blocks:
            x, s = self._single_fwd(blk, x, temb_silu, B, L, cos, sin, save)
            if save:
                sv["single"].append(s)

        # --- norm_out (AdaLayerNormContinuous: chunk order is (scale, shift)) + proj_out, image tokens only
        mod_out, _ = ops.lora_gemv_fwd(temb_silu, m.norm_out.linear.weight, m.norm_out.linear.bias)
        xo = x.view(B, L, D)
        n_out = _empty((B * Li, D), packed)
        mean_o = torch.empty(B * Li, device=dev, dtype=torch.float32)
        rstd_o = torch.empty(B * Li, device=dev, dtype=torch.float32)
        for b in range(B):
            _, mo, ro = ops.ln_modulate_fwd(xo[b, Lt:], mod_out[b:b + 1, D:2 * D], mod_out[b:b + 1, 0:D], Li,
                                            out=n_out[b * Li:(b + 1) * Li])
            mean_o[b * Li:(b + 1) * Li].copy_(mo)
            rstd_o[b * Li:(b + 1) * Li].copy_(ro)
        pred = _empty((B * Li, Cin), packed)
        gemm_bf16(n_out, m.proj_out.weight, pred, bias=m.proj_out.bias)
        if save:
            sv.update(temb_silu=temb_silu, x_final=x, mod_out=mod_out, mean_o=mean_o, rstd_o=rstd_o)
            self.saved = sv
        return pred

    def _double_fwd(self, blk, img, txt, temb_silu, B, Li, Lt, cos, sin, save):
        D, H = self.D, self.H
        L = Lt + Li
        a = blk.attn
        s = {}
        streams = []
        for name, x, norm, qkv_lins, Ls in (("i", img, blk.norm1, (a.to_q, a.to_k, a.to_v), Li),
                                            ("t", txt, blk.norm1_context, (a.add_q_proj, a.add_k_proj, a.add_v_proj), Lt)):
            mod, zmod, amod = self._mod_fwd(norm.linear, temb_silu)
            n, mean, rstd = ops.ln_modulate_fwd(x, mod[:, 0:D], mod[:, D:2 * D], Ls)
            qkv = _empty((B * Ls, 3 * D), x)
            zq = self._qkv_fwd(qkv_lins, n, qkv)
            streams.append(dict(name=name, x=x, mod=mod, zmod=zmod, amod=amod, n=n, mean=mean, rstd=rstd, qkv=qkv, zq=zq, Ls=Ls))
        si, st = streams
        Q = _empty((B, H, L, 128), img)
        K = _empty((B, H, L, 128), img)
        V = _empty((B, H, L, 128), img)
        qt, qi = st["qkv"], si["qkv"]
        ops.qk_norm_rope_fwd(qt[:, :D], qt[:, D:2 * D], qt[:, 2 * D:], a.norm_added_q.weight, a.norm_added_k.weight, cos, sin,
                             Q, K, V, B, Lt, 0)
        ops.qk_norm_rope_fwd(qi[:, :D], qi[:, D:2 * D], qi[:, 2 * D:], a.norm_q.weight, a.norm_k.weight, cos, sin,
                             Q, K, V, B, Li, Lt)
        o_t = _empty((B * Lt, D), img)
        o_i = _empty((B * Li, D), img)
        lse = attention.fwd(Q, K, V, o_t, o_i, Lt)
        outs = []
        for sd, o, out_lin, ff in ((si, o_i, a.to_out[0], blk.ff), (st, o_t, a.to_add_out, blk.ff_context)):
            x, mod, Ls = sd["x"], sd["mod"], sd["Ls"]
            x1 = _empty(x.shape, x)
            y_a = _empty(x.shape, x)
            z_o = self._lin(out_lin, o, x1, gate=mod[:, 2 * D:3 * D], rows_per_sample=Ls, res=x, aux_out=y_a)
            n2, mean2, rstd2 = ops.ln_modulate_fwd(x1, mod[:, 3 * D:4 * D], mod[:, 4 * D:5 * D], Ls)
            inner = ff.net[0].proj.out_features
            pre = _empty((x.shape[0], inner), x)
            act = _empty((x.shape[0], inner), x)
            z_f1 = self._lin(ff.net[0].proj, n2, act, act=ACT_GELU_TANH, aux_out=pre)
            x2 = _empty(x.shape, x)
            y_m = _empty(x.shape, x)
            z_f2 = self._lin(ff.net[2], act, x2, gate=mod[:, 5 * D:6 * D], rows_per_sample=Ls, res=x1, aux_out=y_m)
            outs.append(x2)
            if save:
                sd.update(o=o, x1=x1, y_a=y_a, z_o=z_o, n2=n2, mean2=mean2, rstd2=rstd2, pre=pre, act=act, z_f1=z_f1,
                          z_f2=z_f2, y_m=y_m)
        if save:
            s = dict(i=si, t=st, Q=Q, K=K, V=V, lse=lse)
        return outs[0], outs[1], s

    def _single_fwd(self, blk, x, temb_silu, B, L, cos, sin, save):
        D, H = self.D, self.H
        a = blk.attn
        mod, zmod, amod = self._mod_fwd(blk.norm.linear, temb_silu)
        n, mean, rstd = ops.ln_modulate_fwd(x, mod[:, 0:D], mod[:, D:2 * D], L)
        inner = blk.proj_mlp.out_features
        # one activation buffer per block: [attention out (D) | GELU(mlp) (4D) | q k v pre-norm (3D)]
        big = _empty((B * L, D + inner + 3 * D), x)
        cat = big[:, :D + inner]
        qkv = big[:, D + inner:]
        pre = _empty((B * L, inner), x)
        lins4 = (blk.proj_mlp, a.to_q, a.to_k, a.to_v)
        loras4 = [live_lora(l) for l in lins4]
        grp4 = loras4[0].network_ref().fused_group(loras4) if all(l is not None for l in loras4) else None
        if grp4 is not None:
            zq = ("g4", group_fwd(grp4, lins4, n, big[:, D:], act=ACT_GELU_TANH, act_ncols=inner, aux_out=pre))
            z_mlp = None
        else:
            zq = self._qkv_fwd((a.to_q, a.to_k, a.to_v), n, qkv)
            z_mlp = self._lin(blk.proj_mlp, n, cat[:, D:], act=ACT_GELU_TANH, aux_out=pre)
        Q = _empty((B, H, L, 128), x)
        K = _empty((B, H, L, 128), x)
        V = _empty((B, H, L, 128), x)
        ops.qk_norm_rope_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], a.norm_q.weight, a.norm_k.weight, cos, sin, Q, K, V,
                             B, L, 0)
        lse = attention.fwd(Q, K, V, None, cat[:, :D], 0)
        x1 = _empty(x.shape, x)
        y = _empty(x.shape, x)
        z_out = self._lin(blk.proj_out, cat, x1, gate=mod[:, 2 * D:3 * D], rows_per_sample=L, res=x, aux_out=y)
        s = None
        if save:
            s = dict(x=x, mod=mod, zmod=zmod, amod=amod, n=n, mean=mean, rstd=rstd, qkv=qkv, zq=zq, cat=cat, pre=pre,
                     z_mlp=z_mlp, Q=Q, K=K, V=V, lse=lse, y=y, z_out=z_out)
        return x1, s

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def backward(self, dpred):
        """dpred [B*Li, 64] bf16 -> accumulates dA / dB of every live adapter into the flat gradient buffer."""
        sv = self.saved
        assert sv is not None, "FluxEngine.backward without a saved forward"
        m = self.model
        D = self.D
        B, Li, Lt = sv["B"], sv["Li"], sv["Lt"]
        L = Li + Lt
        cos, sin = sv["cos"], sv["sin"]
        temb_silu = sv["temb_silu"]
        dev = dpred.device
        # proj_out + norm_out (frozen): dx for the image rows of the joint stream, zero for the text rows
        dn = _empty((B * Li, D), dpred)
        gemm_bf16(dpred, m.proj_out.weight, dn, trans_b=True)
        dx = torch.zeros((B * L, D), device=dev, dtype=torch.bfloat16)
        xo = sv["x_final"].view(B, L, D)
        dxv = dx.view(B, L, D)
        mod_out = sv["mod_out"]
        for b in range(B):
            ops.ln_modulate_bwd(dn[b * Li:(b + 1) * Li], xo[b, Lt:], sv["mean_o"][b * Li:(b + 1) * Li],
                                sv["rstd_o"][b * Li:(b + 1) * Li], mod_out[b:b + 1, 0:D], Li, dres=None, out=dxv[b, Lt:])
        for blk, s in zip(reversed(list(m.single_transformer_blocks)), reversed(sv["single"])):
            dx = self._single_bwd(blk, s, dx, temb_silu, B, L, cos, sin)
        d_img = _empty((B * Li, D), dpred)
        d_txt = _empty((B * Lt, D), dpred)
        dxv = dx.view(B, L, D)
        d_txt.view(B, Lt, D).copy_(dxv[:, :Lt])
        d_img.view(B, Li, D).copy_(dxv[:, Lt:])
        n_double = len(m.transformer_blocks)
        for idx, (blk, s) in enumerate(zip(reversed(list(m.transformer_blocks)), reversed(sv["double"]))):
            first = idx == n_double - 1  # block 0: nothing trainable upstream, skip its input gradient
            d_img, d_txt = self._double_bwd(blk, s, d_img, d_txt, temb_silu, B, Li, Lt, cos, sin, need_dx=not first)
        self.saved = None

    def _single_bwd(self, blk, s, dx1, temb_silu, B, L, cos, sin):
        D = self.D
        a = blk.attn
        mod = s["mod"]
        dmod = torch.zeros((B, 3 * D), device=dx1.device, dtype=torch.float32)
        # x1 = x + gate * y
        dy = _empty(dx1.shape, dx1)
        ops.col_reduce(dx1, L, b=s["y"], g=mod[:, 2 * D:3 * D], mul_out=dy, sum_ab=dmod[:, 2 * D:3 * D])
        inner = blk.proj_mlp.out_features
        dbig = _empty((B * L, D + inner + 3 * D), dx1)  # [dO (D) | d mlp pre-activation (4D) | d qkv (3D)]
        dcat = dbig[:, :D + inner]
        dqkv = dbig[:, D + inner:]
        linear_bwd(blk.proj_out, dy, s["cat"], s["z_out"], dcat, lora=live_lora(blk.proj_out),
                   n_slices=[(0, D, {}), (D, D + inner, dict(aux_in=s["pre"]))])
        # attention branch
        dQ, dK, dV = attention.bwd(s["Q"], s["K"], s["V"], None, s["cat"][:, :D], None, dcat[:, :D], s["lse"], 0)
        qkv = s["qkv"]
        ops.qk_norm_rope_bwd(dQ, dK, dV, qkv[:, :D], qkv[:, D:2 * D], a.norm_q.weight, a.norm_k.weight, cos, sin,
                             dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], B, L, 0)
        dn = _empty((B * L, D), dx1)
        if s["zq"][0] == "g4":
            lins4 = (blk.proj_mlp, a.to_q, a.to_k, a.to_v)
            loras4 = [live_lora(l) for l in lins4]
            group_bwd(loras4[0].network_ref().fused_group(loras4), lins4, dbig[:, D:], s["n"], s["zq"][1], dn)
        else:
            # MLP branch: dcat[:, D:] already carries gelu'
            linear_bwd(blk.proj_mlp, dcat[:, D:], s["n"], s["z_mlp"], dn, lora=live_lora(blk.proj_mlp))
            self._qkv_bwd((a.to_q, a.to_k, a.to_v), dqkv, s["n"], s["zq"], dn, accumulate=True)
        # AdaLN: dx = dx1 + dLN(dn); modulation-vector gradients
        dx = ops.ln_modulate_bwd(dn, s["x"], s["mean"], s["rstd"], mod[:, D:2 * D], L, dres=dx1)
        ops.col_reduce(dn, L, b=s["x"], mean=s["mean"], rstd=s["rstd"], sum_a=dmod[:, 0:D], sum_ab=dmod[:, D:2 * D])
        self._mod_bwd(blk.norm.linear, dmod, temb_silu, s["zmod"], s["amod"])
        return dx

    def _double_bwd(self, blk, s, d_img2, d_txt2, temb_silu, B, Li, Lt, cos, sin, need_dx=True):
        D = self.D
        a = blk.attn
        dO = {}
        d1 = {}
        dmods = {}
        for key, dx2, out_lin, ff in (("i", d_img2, a.to_out[0], blk.ff), ("t", d_txt2, a.to_add_out, blk.ff_context)):
            sd = s[key]
            mod, Ls = sd["mod"], sd["Ls"]
            dmod = torch.zeros((B, 6 * D), device=dx2.device, dtype=torch.float32)
            # x2 = x1 + gate_mlp * y_m
            dy = _empty(dx2.shape, dx2)
            ops.col_reduce(dx2, Ls, b=sd["y_m"], g=mod[:, 5 * D:6 * D], mul_out=dy, sum_ab=dmod[:, 5 * D:6 * D])
            dpre = _empty(sd["pre"].shape, dx2)
            linear_bwd(ff.net[2], dy, sd["act"], sd["z_f2"], dpre, lora=live_lora(ff.net[2]), aux_in=sd["pre"])
            dn2 = _empty(dx2.shape, dx2)
            linear_bwd(ff.net[0].proj, dpre, sd["n2"], sd["z_f1"], dn2, lora=live_lora(ff.net[0].proj))
            dx1 = ops.ln_modulate_bwd(dn2, sd["x1"], sd["mean2"], sd["rstd2"], mod[:, 4 * D:5 * D], Ls, dres=dx2)
            ops.col_reduce(dn2, Ls, b=sd["x1"], mean=sd["mean2"], rstd=sd["rstd2"], sum_a=dmod[:, 3 * D:4 * D],
                           sum_ab=dmod[:, 4 * D:5 * D])
            # x1 = x + gate_msa * y_a
            dya = dy  # reuse
            ops.col_reduce(dx1, Ls, b=sd["y_a"], g=mod[:, 2 * D:3 * D], mul_out=dya, sum_ab=dmod[:, 2 * D:3 * D])
            do = _empty(dx2.shape, dx2)
            linear_bwd(out_lin, dya, sd["o"], sd["z_o"], do, lora=live_lora(out_lin))
            dO[key], d1[key], dmods[key] = do, dx1, dmod
        dQ, dK, dV = attention.bwd(s["Q"], s["K"], s["V"], s["t"]["o"], s["i"]["o"], dO["t"], dO["i"], s["lse"], Lt)
        outs = {}
        for key, norm, qkv_lins, wq, wk, off in (("i", blk.norm1, (a.to_q, a.to_k, a.to_v), a.norm_q.weight, a.norm_k.weight, Lt),
                                                 ("t", blk.norm1_context, (a.add_q_proj, a.add_k_proj, a.add_v_proj),
                                                  a.norm_added_q.weight, a.norm_added_k.weight, 0)):
            sd = s[key]
            mod, Ls, dmod = sd["mod"], sd["Ls"], dmods[key]
            qkv = sd["qkv"]
            dqkv = _empty(qkv.shape, qkv)
            ops.qk_norm_rope_bwd(dQ, dK, dV, qkv[:, :D], qkv[:, D:2 * D], wq, wk, cos, sin, dqkv[:, :D], dqkv[:, D:2 * D],
                                 dqkv[:, 2 * D:], B, Ls, off)
            dn = _empty(sd["x"].shape, qkv)
            self._qkv_bwd(qkv_lins, dqkv, sd["n"], sd["zq"], dn, accumulate=False)
            ops.col_reduce(dn, Ls, b=sd["x"], mean=sd["mean"], rstd=sd["rstd"], sum_a=dmod[:, 0:D], sum_ab=dmod[:, D:2 * D])
            outs[key] = ops.ln_modulate_bwd(dn, sd["x"], sd["mean"], sd["rstd"], mod[:, D:2 * D], Ls, dres=d1[key]) \
                if need_dx else None
            self._mod_bwd(norm.linear, dmod, temb_silu, sd["zmod"], sd["amod"])
        return outs["i"], outs["t"]


class FluxFunction(torch.autograd.Function):
    """Autograd seam for eager trainers: forward = FluxEngine.forward, backward = FluxEngine.backward.  The LoRA
    gradients are accumulated in place into the network's flat buffer; no gradient flows to the inputs (cached
    latents / embeddings are data)."""

    @staticmethod
    def forward(ctx, anchor, model, packed, t01, enc, pooled, guidance, txt_ids, img_ids):
        eng = model.engine
        pred = eng.forward(packed, t01, enc, pooled, guidance, txt_ids, img_ids, save=True)
        ctx.engine = eng
        ctx.saved_state = eng.saved
        return pred.view(packed.shape[0], packed.shape[1], -1)

    @staticmethod
    def backward(ctx, dpred):
        eng = ctx.engine
        eng.saved = ctx.saved_state
        eng.backward(dpred.reshape(-1, dpred.shape[-1]).contiguous())
        return (None,) * 9


def flux_apply(model, hidden_states, timestep, encoder_hidden_states, pooled_projections, txt_ids, img_ids, guidance):
    B = hidden_states.shape[0]
    dev = hidden_states.device
    t01 = timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(B).contiguous()
    g = None
    if model.cfg.guidance_embeds:
        if guidance is None:
            raise ValueError("this FLUX variant has guidance embeddings: pass `guidance`")
        g = guidance.to(device=dev, dtype=torch.float32).reshape(-1).expand(B).contiguous()
    if txt_ids.dim() == 3:
        txt_ids = txt_ids[0]
    if img_ids.dim() == 3:
        img_ids = img_ids[0]
    packed = hidden_states.to(torch.bfloat16).contiguous()
    enc = encoder_hidden_states.to(torch.bfloat16).contiguous()
    pooled = pooled_projections.to(torch.bfloat16).contiguous()
    net = model.engine.active_network()
    mods = net.get_all_modules() if net is not None else []
    live = bool(mods) and mods[0].is_live()
    if torch.is_grad_enabled() and live:
        anchor = mods[0].lora_down.weight  # a leaf that requires grad, so that autograd calls our backward
        return FluxFunction.apply(anchor, model, packed, t01, enc, pooled, g, txt_ids, img_ids)
    pred = model.engine.forward(packed, t01, enc, pooled, g, txt_ids, img_ids, save=False)
    return pred.view(B, packed.shape[1], -1)
