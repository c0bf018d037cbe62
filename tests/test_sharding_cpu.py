"""Host-side logic of the N > 1 path on CPU: contiguous batch sharding and the single
all_gather of logits + beta, exercised with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_the_batch_exactly():
    from hriemo.pipeline import shard_bounds

    for B in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))   # contiguous, no overlap
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1                                     # balanced


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, n_e, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hriemo.pipeline import gather_outputs, shard_bounds

        g = torch.Generator().manual_seed(123)               # every rank builds the same "global" result
        logits_all = torch.randn(B, n_e, generator=g)
        beta_all = torch.rand(B, 1, generator=g)
        lo, hi = shard_bounds(B, rank, world)
        logits, beta = gather_outputs(logits_all[lo:hi].clone(), beta_all[lo:hi].clone())
        ok = torch.equal(logits, logits_all) and torch.equal(beta, beta_all)
        out_q.put((rank, bool(ok), tuple(logits.shape), tuple(beta.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_outputs_world_size_2_gloo():
    world, B, n_e = 2, 12, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, n_e, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res == [(0, True, (B, n_e), (B, 1)), (1, True, (B, n_e), (B, 1))]


def test_h2d_byte_accounting_matches_the_staging_schedule():
    """pipeline.h2d_bytes mirrors forward_from_host's slab schedule: every second slab travels as bf16."""
    from hriemo import pipeline

    per = 564 * 768 * 4
    assert pipeline.h2d_bytes(4096, per, slab=512, host_cast_every=0) == 4096 * per
    assert pipeline.h2d_bytes(4096, per, slab=512, host_cast_every=2) == 4096 * per * 3 // 4   # 4 of 8 slabs halved
    # ramp 64+128+256 (fp32), then 512-slabs alternating fp32 / host-cast, 64 left over (host-cast)
    assert pipeline.h2d_bytes(4096, per, slab=512, host_cast_every=2, ramp=True) == (4096 - 1600) * per + 1600 * per // 2
    assert pipeline.h2d_bytes(1000, per, slab=512, host_cast_every=2) == 1000 * per            # two slabs: no pre-cast
    assert pipeline.h2d_bytes(1100, per, slab=512, host_cast_every=2) == (512 + 76) * per + 512 * per // 2


def test_host_cast_default_follows_the_rank_thread_budget():
    from hriemo import pipeline

    assert pipeline.default_host_cast_every(16) == 2 and pipeline.default_host_cast_every(8) == 2
    assert pipeline.default_host_cast_every(4) == 0 and pipeline.default_host_cast_every(1) == 0


def test_slab_schedule_covers_the_batch_in_order():
    """Every plan is a partition of [0, B) in order, no slab exceeds the staging size, ramp slabs are
    never host-cast, and small batches get no ramp."""
    from hriemo import pipeline

    for B in (1, 3, 511, 512, 1000, 1100, 2048, 2559, 2560, 4096, 16384):
        for slab in (2, 64, 512):
            for every in (0, 2, 3):
                plan = pipeline.slab_schedule(B, slab, every, ramp=(B % 2 == 0))
                assert plan[0][0] == 0 and plan[-1][1] == B
                assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
                assert all(0 < e - s <= slab for s, e, _ in plan)
                if every == 0:
                    assert not any(h for _, _, h in plan)
    plan = pipeline.slab_schedule(4096, 512, 2, ramp=True)
    assert [e - s for s, e, _ in plan[:3]] == [64, 128, 256] and not any(h for _, _, h in plan[:4])
    assert [e - s for s, e, _ in pipeline.slab_schedule(2048, 512, 2)] == [512] * 4


def _load_bench():
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    return bench


def test_bench_deadline_prints_the_line_when_a_supplementary_leg_hangs():
    """bench.py's supplementary legs (training step over NCCL) run under a deadline: when one hangs, the contract
    line assembled so far is emitted with the leg marked failed and the process exits 0.  Exercised in a child."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, time, json, importlib.util\n"
        f"spec = importlib.util.spec_from_file_location('b', r'{os.path.join(root, 'bench.py')}')\n"
        "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
        "line = {'metric': 'm', 'value': 1.0}\n"
        "def emit(extra=None):\n"
        "    out = dict(line); out.update(extra or {}); print(json.dumps(out), flush=True)\n"
        "with b.Deadline(0.3, emit, 'train_step'):\n"
        "    time.sleep(30)\n"
        "print('NOT REACHED')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "NOT REACHED" not in r.stdout
    import json
    rec = json.loads(r.stdout.strip().splitlines()[-1])
    assert rec["value"] == 1.0 and "error" in rec["train_step"]


def test_bench_parity_leg_statistics():
    """The parity block of the bench line, on the CPU with the oracle itself standing in for the GPU model (plus a
    known perturbation): denominators, excluded counts and agreement rates are what the definitions say."""
    import hriemo_oracle as O
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    bench = _load_bench()
    torch.manual_seed(1234)
    net = FusionWithEmotionDecoder(d_model=64, n_heads=2, beta_hidden=16, num_layers_fusion=1, num_layers_decoder=1).eval()
    sd = {k: v.detach() for k, v in net.state_dict().items()}

    class Standin:
        def state_dict(self):
            return sd

        def __call__(self, h_a, h_t, m_a, m_t):
            lo, be, z = O.fusion_with_emotion_decoder(sd, h_a, h_t, m_a, m_t, n_heads=2)
            lo = lo.clone()
            lo[0, 0] = -lo[0, 0] if lo[0, 0].abs() > 0.05 else lo[0, 0] + 1.0     # one wrong decision per chunk, outside tol
            return lo, be, z

    parity, cpu = bench.parity_and_cpu_leg(Standin(), torch.device("cpu"), 12, 6, 64, 64, 2, False, 32, 8)
    assert parity["n"] == 32 and parity["n_ragged"] == 16 and parity["decisions"] == 128
    assert parity["thr_disagreements"] == 4 and parity["thr_disagreements_outside_tol"] == 4
    assert abs(parity["thr_agree_all"] - 124 / 128) < 1e-12
    assert parity["thr_agree"] is not None and parity["thr_agree"] < 1.0 and parity["thr_excluded"] >= 0
    assert parity["beta_max_abs"] == 0.0 and parity["beta_gt_half_agree_all"] == 1.0 and parity["beta_batch_argmax_equal"]
    assert cpu["kind"] == "port" and cpu["value"] > 0 and cpu["cores"] >= 1


@pytest.mark.parametrize("n_slabs,t_copy,t_pack", [(8, 0.004, 0.002), (8, 0.002, 0.02), (3, 0.002, 0.001), (1, 0.001, 0.001),
                                                   (16, 0.003, 0.004)])
def test_two_ended_plan_hands_every_slab_out_once(n_slabs, t_copy, t_pack):
    """forward_from_host's default plan: the copy side takes slabs from the front, the host side converts from the
    back; whatever their speeds, every slab is handed out exactly once, converted slabs are sent under the index the
    host side gave them, and the batch does not end with a long wait for the host."""
    import threading
    import time

    from hriemo import pipeline

    plan = pipeline.TwoEndedPlan(n_slabs)
    packed = []

    def host():
        k = 0
        while True:
            i = plan.claim_back()
            if i is None:
                break
            time.sleep(t_pack)
            packed.append((i, k))
            plan.publish(i, k, t_pack)
            k += 1
        plan.host_done()

    th = threading.Thread(target=host)
    th.start()
    sent = []
    t_last_front = None
    while True:
        item = plan.next()
        if item is None:
            break
        sent.append(item)
        time.sleep(t_copy if item[1] < 0 else t_copy / 2)
    th.join()
    assert sorted(i for i, _ in sent) == list(range(n_slabs))
    assert sorted(x for x in sent if x[1] >= 0) == sorted(packed)
    fronts = [i for i, k in sent if k < 0]
    assert fronts == sorted(fronts) and (not fronts or fronts[0] == 0)
    if n_slabs >= 8 and t_pack <= t_copy:
        assert len(packed) >= n_slabs // 2 - 1      # a fast host takes about half or more
    if t_pack >= 5 * t_copy:
        assert len(packed) <= 2                     # a slow host only what it can finish in time


def test_two_ended_plan_ignores_unpaced_decisions_and_keeps_its_rates():
    """The first two decisions of a call only fill the copy queue (0.2 ms apart): measured as the copy side's rate they
    made the host side stop after one slab on every later call of a stream.  Unpaced intervals are not measured, and a
    plan can start from the rates of the previous call."""
    from hriemo import pipeline

    plan = pipeline.TwoEndedPlan(8)
    assert plan.claim_back() == 7
    assert plan.next(paced=False) == (0, -1) and plan.next(paced=False) == (1, -1)
    assert plan.t_step is None
    plan.publish(7, 0, 0.017)
    assert plan.claim_back() == 6                    # rates unknown: keeps going while more than two are left
    assert plan.next(paced=True) == (7, 0) and plan.t_step is None   # the interval in front of the first paced one: no
    import time
    time.sleep(0.01)
    assert plan.next(paced=True) == (2, -1) and plan.t_step >= 0.01

    plan = pipeline.TwoEndedPlan(8, t_pack=0.017, t_step=0.012)
    assert [plan.claim_back() for _ in range(3)] == [7, 6, 5]
    slow = pipeline.TwoEndedPlan(8, t_pack=0.2, t_step=0.012)      # a slow host (2 threads): nothing is worth converting
    assert slow.claim_back() is None


def test_two_ended_plan_surfaces_a_host_failure():
    from hriemo import pipeline

    plan = pipeline.TwoEndedPlan(4)
    assert plan.claim_back() == 3
    plan.host_done(RuntimeError("pack failed"))
    with pytest.raises(RuntimeError, match="pack failed"):
        plan.next()
