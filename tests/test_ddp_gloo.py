"""N > 1 semantics on CPU (gloo, world_size 2): the flat LoRA gradient buffer is summed over ranks and the 1/W of the
average is folded into the clip/AdamW prescale (train_step._all_reduce, optimizer hyper[7]); W ranks with one sample
each == one rank accumulating W samples and dividing by W.  The gradients here come from the
oracle on CPU; the kernels themselves are covered by the gpu tests."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flux_ref, lora_ref, make_golden


def _grads_for_sample(model, net, batch, idx):
    net.zero_grad(set_to_none=True)
    sl = slice(idx, idx + 1)
    noisy = lora_ref.add_noise_flowmatch(batch["latents"][sl], batch["noise"][sl], batch["timesteps"][sl])
    with net:
        pred = lora_ref.flux_predict(model, noisy, batch["timesteps"][sl], batch["text"][sl], batch["pooled"][sl], 1.0,
                                     flux_ref.pack_latents, flux_ref.unpack_latents, flux_ref.make_img_ids)
        lora_ref.flow_loss(pred, batch["latents"][sl], batch["noise"][sl]).backward()
    return torch.cat([p.grad.reshape(-1) for l in net.loras for p in (l.lora_down.weight, l.lora_up.weight)])


def _build():
    cfg, model, batch = make_golden.build(seed=11)
    torch.manual_seed(3)
    net = lora_ref.LoRANetworkRef(model, lora_dim=4)
    with torch.no_grad():
        for l in net.loras:
            l.lora_up.weight.normal_(0, 0.05)
    return model, net, batch


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model, net, batch = _build()
    flat = _grads_for_sample(model, net, batch, rank)  # rank k draws sample k (batch sharding, section 8e)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)        # what FluxLoRATrainStep._all_reduce does on flat_grads
    flat = flat * (1.0 / world)                         # hyper[7] = grad_prescale = 1 / world inside the AdamW kernel
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_two_ranks_equal_accumulation(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, 29581, out), nprocs=2, join=True)
    got = torch.load(out)
    model, net, batch = _build()
    want = (_grads_for_sample(model, net, batch, 0) + _grads_for_sample(model, net, batch, 1)) / 2
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-8)


# ---------------------------------------------------------------------------------------------------------------------
# the PRODUCT's N > 1 plumbing: FluxLoRATrainStep.__init__ (parameter broadcast, grad_prescale = 1/W) and ._all_reduce()
# ---------------------------------------------------------------------------------------------------------------------
def _product_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from ai_toolkit_b200 import LoRASpecialNetwork
    from ai_toolkit_b200.flux import FluxConfig, FluxTransformer2DModel
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep

    cfg, omodel, batch = make_golden.build(seed=11)
    torch.manual_seed(100 + rank)  # every process draws its OWN lora_down (kaiming) / lora_up values
    model = FluxTransformer2DModel(FluxConfig(num_layers=cfg.num_layers, num_single_layers=cfg.num_single_layers,
                                              num_attention_heads=cfg.num_attention_heads,
                                              joint_attention_dim=cfg.joint_attention_dim,
                                              pooled_projection_dim=cfg.pooled_projection_dim), dtype=torch.float32)
    model.load_state_dict(omodel.state_dict())
    net = LoRASpecialNetwork(None, model, lora_dim=4, alpha=4, train_text_encoder=False, is_flux=True, transformer_only=True)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    with torch.no_grad():
        for l in net.get_all_modules():
            l.lora_up.weight.normal_(0, 0.05)
    mine_before = net.flat_params.clone()
    opt = B200AdamW(net, lr=1e-4, ema_decay=0.99)  # grad_prescale left at its default 1.0 on purpose
    step = FluxLoRATrainStep(model, net, opt, batch_size=1, latent_shape=(16, 8, 8), text_len=8, use_cuda_graph=False)
    # (1) replicas start from rank 0's values; the EMA shadow follows; the packs are marked stale
    gathered = [torch.empty_like(net.flat_params) for _ in range(world)]
    dist.all_gather(gathered, net.flat_params)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    assert rank == 0 or not torch.equal(mine_before, net.flat_params)
    assert torch.equal(opt.ema, net.flat_params) and net._pack_dirty
    # (2) the average's 1/W is in the optimizer's device hyper-parameters without the caller doing anything
    assert abs(float(opt.hyper[7]) - 1.0 / world) < 1e-7 and opt.grad_prescale == 1.0 / world
    # (3) gradients of this rank's sample (oracle on the broadcast adapter values) through the product's all-reduce
    onet = lora_ref.LoRANetworkRef(omodel, lora_dim=4)
    with torch.no_grad():
        for a, b in zip(net.get_all_modules(), onet.loras):
            assert a.lora_name == b.lora_name
            b.lora_down.weight.copy_(a.lora_down.weight)
            b.lora_up.weight.copy_(a.lora_up.weight)
    g = _grads_for_sample(omodel, onet, batch, rank)
    net.ensure_grad_views()
    net.flat_grads[:g.numel()].copy_(g)
    step._all_reduce()
    eff = net.flat_grads[:g.numel()] * opt.hyper[7]  # what clip_adamw_kernel sees (grad * hyper[7])
    if rank == 0:
        torch.save({"eff": eff.clone(), "params": net.flat_params.clone()}, out)
    dist.destroy_process_group()


def test_product_train_step_two_ranks(tmp_path):
    """FluxLoRATrainStep under world_size 2 (gloo): rank-0 parameters everywhere, prescale 1/W set by the class, and
    all-reduce + prescale == accumulation over the W samples / W on the product's own buffers."""
    out = str(tmp_path / "p.pt")
    mp.spawn(_product_worker, args=(2, 29583, out), nprocs=2, join=True)
    got = torch.load(out)
    from ai_toolkit_b200 import LoRASpecialNetwork  # noqa: F401  (same import side effects as the workers)
    cfg, omodel, batch = make_golden.build(seed=11)
    onet = lora_ref.LoRANetworkRef(omodel, lora_dim=4)
    off = 0
    with torch.no_grad():
        for l in onet.loras:
            for p in (l.lora_down.weight, l.lora_up.weight):
                p.copy_(got["params"][off:off + p.numel()].view(p.shape))
                off += p.numel()
    want = (_grads_for_sample(omodel, onet, batch, 0) + _grads_for_sample(omodel, onet, batch, 1)) / 2
    torch.testing.assert_close(got["eff"], want, rtol=1e-5, atol=1e-8)
