"""Data-parallel training step, world size 2 over gloo on the CPU: each rank runs hriemo.train.Trainer on its shard
of the utterances (float64 kernel stand-ins, tests/kernel_standins.py) and the gradient arena is averaged with ONE
all-reduce; the updated parameters must equal those of a single process stepping on the whole batch (the loss is a
mean over utterances and the shards are equal-sized, so the mean of the shard gradients is the batch gradient) and
must be identical on both ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _setup(patch, B, T_a, T_t, d, H, Ne):
    import kernel_standins
    kernel_standins.install(patch, exact=True)
    from hriemo import backward
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    patch.setattr(backward.E, "to_seq", lambda x, what, ld=None: backward.E.Seq(x.reshape(-1, x.shape[-1]), x.shape[0], x.shape[1]))
    torch.manual_seed(31)
    model = FusionWithEmotionDecoder(d_model=d, num_emotions=Ne, n_heads=H, num_layers_fusion=1, num_layers_decoder=1,
                                     beta_hidden=32, dropout=0.0).double()
    g = torch.Generator().manual_seed(32)
    h_a = torch.randn(B, T_a, d, generator=g, dtype=torch.float64)
    h_t = torch.randn(B, T_t, d, generator=g, dtype=torch.float64)
    labels = torch.eye(Ne, dtype=torch.float64)[torch.randint(0, Ne, (B,), generator=g)]
    return model, h_a, h_t, labels


def _worker(rank, world, port, shape, out_q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.join(os.path.dirname(here), "hri-emo_b200"), os.path.join(os.path.dirname(here), "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hriemo.pipeline import shard_bounds
        from hriemo.train import Trainer

        model, h_a, h_t, labels = _setup(_Patch(), *shape)
        lo, hi = shard_bounds(shape[0], rank, world)
        if rank == 1:
            # a replica that was built differently (other seed / checkpoint): the Trainer must start it from rank 0's
            # parameters (DistributedDataParallel broadcasts module state at construction), or the ranks diverge
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(0.25)
        trainer = Trainer(model, lr=1e-3, max_norm=0.5)
        assert trainer.distributed
        for _ in range(2):
            info = trainer.step(h_a[lo:hi], h_t[lo:hi], None, None, labels[lo:hi])
        out_q.put((rank, trainer.params.numpy().copy(), float(info["grad_norm"])))   # by value: the worker may exit first
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_trainer_world_size_2_gloo_equals_single_process(monkeypatch):
    shape = (4, 10, 6, 128, 2, 4)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert (res[0][1] == res[1][1]).all()              # replicas stay bit-identical
    assert res[0][2] == res[1][2]

    from hriemo.train import Trainer
    model, h_a, h_t, labels = _setup(monkeypatch, *shape)   # single-process arm, same stand-ins (undone after the test)
    single = Trainer(model, lr=1e-3, max_norm=0.5, distributed=False)
    for _ in range(2):
        info = single.step(h_a, h_t, None, None, labels)
    assert (single.params - torch.from_numpy(res[0][1])).abs().max().item() <= 1e-12
    assert abs(float(info["grad_norm"]) - res[0][2]) <= 1e-12
