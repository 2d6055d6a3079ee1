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
