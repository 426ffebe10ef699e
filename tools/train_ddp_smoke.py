"""Data-parallel FS-EEND training through the drop-in: torch DistributedDataParallel (NCCL) around the model whose
forward / backward run in this library's kernels.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_ddp_smoke.py
Each rank takes its own synthetic batch; after backward every rank must hold the same (averaged) gradients, and the step
time is the max over ranks.  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    from fseend_b200.loss import standard_loss
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    torch.manual_seed(0)                                     # same initial weights on every rank
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                       dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().train()
    ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.Adam(ddp.parameters(), lr=1e-5)
    B, T, S = 32, 500, 4
    g = torch.Generator(device="cuda").manual_seed(100 + rank)   # different data per rank
    src = [torch.randn(T, 345, device="cuda", generator=g) for _ in range(B)]
    tgt = [(torch.rand(T, S, device="cuda", generator=g) < 0.3).float() for _ in range(B)]
    lab = [torch.cat([1 - t.max(-1, keepdim=True)[0], t, torch.zeros(T, 1, device="cuda")], -1) for t in tgt]
    lens = [T] * B
    losses, times = [], []
    for step in range(6):
        torch.manual_seed(1000 + step * world + rank)        # dropout streams differ per rank and step
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad(set_to_none=True)
        out, emb_loss, _, _ = ddp(src, lab, lens)
        loss = standard_loss(out, lab) + emb_loss
        loss.backward()
        if step == 0:
            # DDP averaged the gradients: every rank must hold the same values
            gsum = torch.stack([p.grad.double().abs().sum() for p in m.parameters() if p.grad is not None]).sum()
            lo, hi = gsum.clone(), gsum.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool((hi - lo).abs() <= 1e-12 * hi.abs())
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        l = loss.detach().clone()
        dist.all_reduce(l)
        losses.append(l.item() / world)
        times.append(t.item())
    # weights stay in sync across ranks
    wsum = torch.stack([p.detach().double().sum() for p in m.parameters()]).sum()
    lo, hi = wsum.clone(), wsum.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"what": "FS-EEND DDP training through the drop-in (NCCL gradient all-reduce by torch DDP)",
                          "n_gpus": world, "per_gpu_batch": B, "frames": T, "dropout": 0.1,
                          "grads_identical_across_ranks": same, "weights_in_sync": bool((hi - lo).abs() <= 1e-9 * hi.abs()),
                          "mean_loss_per_step": [round(x, 5) for x in losses],
                          "ms_per_step_max_over_ranks": [round(x, 1) for x in times],
                          "frames_per_s_last_step": round(world * B * T / times[-1] * 1e3)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
