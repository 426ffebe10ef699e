"""world_size-2 gloo test (CPU) of the multi-process host logic: sequence sharding and max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fseend_b200.parallel import max_over_ranks, shard_range, shard_sequences


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n and (b - a) - n // world in (0, 1)
                got += list(range(a, b))
            assert got == list(range(n))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = list(range(11))
    mine = shard_sequences(items, rank, world)
    # each rank "times" its shard; the job time is the slowest rank's
    t = max_over_ranks([float(len(mine)), 10.0 - rank], torch.device("cpu"))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((t, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_gloo_sharding_and_max_timing():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert t == [6.0, 10.0]
    assert gathered[0] + gathered[1] == list(range(11))


def _train_worker(rank, world, port, out):
    """One data-parallel training step of the drop-in model on CPU: the train graph with torch stand-ins for the kernels
    (tests/test_train_graph_cpu.py) under DistributedDataParallel over gloo."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fseend_b200.autograd as A
    import fseend_b200.train_graph as G
    import test_train_graph_cpu as S
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    from oracle import fs_eend_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for mod in (A, G):
        mod.LinearFn, mod.AddLayerNormFn = S._Lin, S._AddLn
    A.FfnFn, A.CausalAttnFn, A.SpeakerAttnFn = S._Ffn, S._Causal, S._Spk
    G.L2NormFn, G.HeadFn, G.batch_norm_forward, G._require_device = S._L2, S._Head, S._bn, (lambda dev: None)
    torch.set_num_threads(2)
    sd = O.random_state_dict(seed=3, enc_n_layers=1, dec_n_layers=1)
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=1, dec_n_layers=1,
                                       dropout=0.0, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048)
    m.load_state_dict(sd)
    m = m.double().train()
    m.enc.bn.eval()                      # frozen statistics: the two-rank average then equals the single-process batch mean
    import contextlib
    m._on_device = contextlib.nullcontext    # the wrapper pins the CUDA device of its parameters; there is none here
    ddp = torch.nn.parallel.DistributedDataParallel(m, find_unused_parameters=True)
    lens_all, spk_all = [40, 31, 36, 28], [3, 2, 3, 3]
    src_all, _ = O.synthetic_features(4, 40, seed=5, lens=lens_all)
    tgt_all = [t.double() for t in O.synthetic_labels(5, lens_all, spk_all)]
    sel = [2 * rank, 2 * rank + 1]       # each rank takes its own two recordings
    src, tgt, lens = [src_all[i].double() for i in sel], [tgt_all[i] for i in sel], [lens_all[i] for i in sel]
    o, el, _, _ = ddp(src, tgt, lens)
    # per-rank loss as a SUM over frames so that DDP's gradient average is the full-batch gradient / world
    loss = sum(torch.nn.functional.binary_cross_entropy_with_logits(y, t, reduction="sum") for y, t in zip(o, tgt))
    loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])
    gathered = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    if rank == 0:
        # single-process reference over all four recordings
        ref = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=1, dec_n_layers=1,
                                             dropout=0.0, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048)
        ref.load_state_dict(sd)
        ref = ref.double().train()
        ref.enc.bn.eval()
        total = 0.0
        for r in range(world):           # rank-sized calls: the padded length (and the emb-loss mean) is per call
            idx = [2 * r, 2 * r + 1]
            oo, _, _, _ = G.fs_forward_train(ref, [src_all[i].double() for i in idx], [tgt_all[i] for i in idx],
                                             [lens_all[i] for i in idx])
            total = total + sum(torch.nn.functional.binary_cross_entropy_with_logits(y, tgt_all[i], reduction="sum")
                                for y, i in zip(oo, idx))
        (total / world).backward()
        gref = torch.cat([p.grad.reshape(-1) for p in ref.parameters() if p.grad is not None])
        out.put((bool(torch.equal(gathered[0], gathered[1])), float((gathered[0] - gref).abs().max()), float(gref.abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_gloo_data_parallel_training_step():
    """N > 1 training host logic (bench.py --train, tools/train_ddp_smoke.py): DistributedDataParallel around the drop-in
    model leaves both ranks with identical gradients, equal to the average of the per-rank gradients computed in one
    process."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    import time
    deadline = time.time() + 300
    res = None
    while res is None and time.time() < deadline:
        try:
            res = out.get(timeout=2)
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    if res is None:
        for p in procs:
            p.kill()
        raise AssertionError("data-parallel workers failed")
    same, err, scale = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same
    assert err <= 1e-10 * scale
