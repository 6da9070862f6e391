"""World-size-2 gloo test (CPU) of the data-parallel host logic bench.py uses at N > 1: utterances are sharded by
rank (no data-path collective), gradients are averaged by ONE flat all-reduce (mamba_asr_b200.dist_utils), and the
reduced gradient equals the gradient of the same global batch on one process.  The CPU arm swaps the sm_100a mixers for the oracle (tests may use oracle/)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build():
    from mamba_asr_b200.encoder import ConMambaCTC
    from oracle.cpu_encoder import to_cpu_reference
    torch.manual_seed(0)
    m = ConMambaCTC(d_model=32, d_ffn=64, num_layers=1, n_fft=400, win_length=25, n_mels=80, output_neurons=11, dropout=0.0)
    return to_cpu_reference(m, 400, 80, 25)


def _loss(model, wav, tgt):
    logp = model(wav)
    Bt, L, _ = logp.shape
    return F.ctc_loss(logp.transpose(0, 1), tgt, torch.full((Bt,), L), torch.full((Bt,), tgt.shape[1]), blank=0,
                      reduction="sum", zero_infinity=True)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model = _build()
    g = torch.Generator().manual_seed(1)
    wav = 0.1 * torch.randn(4, 3200, generator=g)
    tgt = torch.randint(1, 11, (4, 3), generator=g)
    shard = slice(rank * 2, rank * 2 + 2)                      # utterance sharding, 2 per rank
    # InputNormalization uses batch statistics: give every rank the same statistics by normalising features outside
    feats = model.features(wav)[shard]

    from mamba_asr_b200.dist_utils import allreduce_gradients
    out = F.log_softmax(model.ctc_lin(model.encode(feats)), dim=-1)
    loss = F.ctc_loss(out.transpose(0, 1), tgt[shard], torch.full((2,), out.shape[1]), torch.full((2,), 3), blank=0,
                      reduction="sum", zero_infinity=True)
    loss.backward()
    local = [None if p.grad is None else p.grad.clone() for p in model.parameters()]
    allreduce_gradients(model.parameters(), world)             # one flat all-reduce over gloo, averaged
    grads = [p.grad.clone() * world for p in model.parameters() if p.grad is not None]
    # the training step bench.py runs at N > 1: all-reduce + clip + AdamW + Noam through TrainStep (torch path on CPU)
    from mamba_asr_b200.trainer import TrainStep
    for p, g_ in zip(model.parameters(), local):
        p.grad = g_
    ts = TrainStep(model, lr=1e-2, max_grad_norm=5.0, n_warmup_steps=10, world_size=world)
    ts.step(dist)
    stepped = [p.detach().clone() for p in model.parameters()]
    t = torch.tensor([float(loss.detach())])
    dist.all_reduce(t)
    if rank == 0:
        q.put((float(t), [g_.numpy().copy() for g_ in grads], [s_.numpy().copy() for s_ in stepped]))   # by value
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradients_equal_single_process_global_batch():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, grads, stepped = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.set_num_threads(2)
    model = _build()
    g = torch.Generator().manual_seed(1)
    wav = 0.1 * torch.randn(4, 3200, generator=g)
    tgt = torch.randint(1, 11, (4, 3), generator=g)
    feats = model.features(wav)
    out = F.log_softmax(model.ctc_lin(model.encode(feats)), dim=-1)
    loss = F.ctc_loss(out.transpose(0, 1), tgt, torch.full((4,), out.shape[1]), torch.full((4,), 3), blank=0,
                      reduction="sum", zero_infinity=True)
    loss.backward()
    ref = [p.grad for p in model.parameters() if p.grad is not None]
    assert abs(float(loss) - total) <= 1e-4 * abs(float(loss))
    assert len(ref) == len(grads)
    for a, b in zip(grads, ref):
        torch.testing.assert_close(torch.from_numpy(a), b, rtol=1e-4, atol=1e-5)
    # one optimizer step on the averaged gradient = what every rank ended with
    from mamba_asr_b200.trainer import TrainStep
    for p in model.parameters():
        if p.grad is not None:
            p.grad.mul_(0.5)
    TrainStep(model, lr=1e-2, max_grad_norm=5.0, n_warmup_steps=10).step()
    for a, p in zip(stepped, model.parameters()):
        torch.testing.assert_close(torch.from_numpy(a), p.detach(), rtol=1e-4, atol=1e-6)
