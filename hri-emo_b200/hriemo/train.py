"""The training step of the reference (BASELINE config 5) as a schedule of C-ABI kernels.

One iteration of train_one_epoch (scripts/fusion/train_fusion_seq_level_decoder.py:300-339, setup :405-416):

    logits, beta, _ = model(h_a, h_t, m_a, m_t)                  :311      backward.loss_and_gradients (forward half)
    loss = BCEWithLogitsLoss()(logits, labels)                   :319      hriemo_bce_beta_loss
    loss = loss - 0.01 * (beta * (1 - beta)).mean()              :326-327       "
    loss.backward()                                              :332      backward.loss_and_gradients (backward half)
    clip_grad_norm_(model.parameters(), max_norm=5.0)            :333      hriemo_grad_norm_clip
    optimizer.step(); optimizer.zero_grad()                      :334-335  hriemo_adamw_step

Layout: every parameter of the model is a view into ONE flat fp32 arena (and so are the gradients and the two AdamW
moments), in `named_parameters()` order with each tensor starting on a 16-byte boundary.  The global gradient norm is
then one reduction, the update one launch, and the data-parallel exchange (one process per GPU, utterances sharded,
SURVEY sec. 8e) ONE all-reduce of the gradient arena over NCCL — 218 MB for the default model — instead of 119.

Dropout is applied when the model was built with p > 0 and is in train() mode (hriemo/dropout.py: counter-based masks
regenerated in the backward; the parity configuration of SURVEY sec. 8d config 5 is dropout = 0).  No autograd graph is
involved; torch provides device memory, streams and the process group.
"""
from __future__ import annotations

import warnings
from typing import Dict, Optional

import torch

from . import backward
from . import engine as E
from . import lib as L
from . import ops

f32 = torch.float32


def invalidate_prepared(model: torch.nn.Module) -> None:
    """The bf16 / fused operand caches (engine.Prepared) key on torch's tensor version counters, which a C kernel
    writing through a raw pointer does not bump: after an optimizer step they are dropped explicitly."""
    for m in model.modules():
        prep = getattr(m, "_prep", None)
        if prep is not None:
            prep._key = None
            prep._value = None


class Trainer:
    """AdamW training of a FusionWithEmotionDecoder (or MOSEI wrapper's inner model) on the B200 path.

    trainer = Trainer(model)                      # model on a CUDA device; its parameters move into the flat arena
    info = trainer.step(h_a, h_t, m_a, m_t, y)    # -> {"loss", "grad_norm", "clip", "logits", "beta"} (device tensors)

    process_group / torch.distributed initialised: DistributedDataParallel's semantics -- rank 0's parameters are
    broadcast when the arena is built (after checking that every rank holds the same layout), and gradients are
    averaged over the ranks with one all-reduce of the arena before the clip.

    graph=True: from the third step with the same input shapes on, the forward + backward + arena fill (~500 kernel
    launches, which at the reference's batch sizes take longer to enqueue from Python than to run) are replayed from
    one CUDA graph; the all-reduce, the clip and AdamW (whose bias corrections depend on the step number) stay eager.
    Inputs are copied into the graph's static buffers; the returned tensors are overwritten by the next step."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, weight_decay: float = 1e-2, betas=(0.9, 0.999),
                 eps: float = 1e-8, max_norm: float = 5.0, beta_weight: float = 0.01, process_group=None,
                 distributed: Optional[bool] = None, graph: bool = False, overlap: Optional[bool] = None):
        p_drop = max((float(getattr(m, "p_drop", 0.0) or 0.0) for m in model.modules()), default=0.0)
        if p_drop > 0 and graph:
            warnings.warn(f"hri-emo_b200 Trainer: the model was built with dropout={p_drop}; the dropout masks are functions of "
                          "per-step keys that a captured graph would freeze, so steps taken in train() mode run eagerly "
                          "instead of through CUDA-graph replay", stacklevel=2)
        self.p_drop = p_drop
        params = list(model.named_parameters())
        if not params:
            raise L.HriemoError("Trainer: the model has no parameters")
        dev = params[0][1].device
        E.require_cuda(params[0][1], "Trainer: model parameter")
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.max_norm, self.beta_weight = max_norm, beta_weight
        self.process_group = process_group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = distributed
        self.slots: Dict[str, tuple] = {}
        off = 0
        for name, p in params:
            if p.dtype != f32 or p.device != dev:
                raise L.HriemoError(f"Trainer: parameter {name} must be fp32 on {dev}")
            self.slots[name] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4          # every tensor starts on a 16-byte boundary
        self.numel = off
        self.params = torch.zeros(off, dtype=f32, device=dev)
        self.grads = torch.zeros(off, dtype=f32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=f32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=f32, device=dev)
        with torch.no_grad():
            for name, p in params:
                o, n = self.slots[name]
                view = self.params[o:o + n].view(p.shape)
                view.copy_(p.data)
                p.data = view                          # the module's tensors ARE the arena from here on
        if self.distributed:
            self._sync_initial_state()
        # Overlapped exchange (data parallel, NCCL): the backward reaches encoder layer 0 last, so the gradients of every
        # other parameter -- the tail of the arena in named_parameters() order: layers >= 1, gate, decoder, 65 % of the
        # bytes at the default model -- are all-reduced on NCCL's stream while layer 0's backward runs; only layer 0's
        # share is exchanged after the backward.  (Trainer(graph=True) replays the step as two graphs around the launch.)
        self._early_names, self._tail_off = self._early_layout()
        self.overlap = (True if overlap is None else bool(overlap)) and self._early_names is not None
        self._early_work = None
        self._reduced = False
        self.allreduce_events = None                   # a list -> (start, end) CUDA events of every all-reduce (benchmarks)
        self.step_count = 0
        self.use_graph = graph
        self._graph = None
        self._graph_key = None
        self._graph_seen = 0
        self._static_in = None
        self._graph_out = None
        invalidate_prepared(model)

    def _sync_initial_state(self) -> None:
        """DistributedDataParallel's contract at construction: every rank starts from rank 0's parameters.  Only the
        gradients are exchanged afterwards, so replicas built from different seeds or checkpoints would otherwise
        train apart silently.  The arena layout (names, sizes) must be the same on every rank: checked first.
        The AdamW moments start at zero everywhere; on a resume, restore `exp_avg`, `exp_avg_sq` and `step_count`
        identically on every rank (they are never exchanged)."""
        import torch.distributed as dist

        world = dist.get_world_size(self.process_group)
        if world == 1:
            return
        sizes = torch.tensor([self.numel, len(self.slots)], dtype=torch.int64, device=self.params.device)
        lo, hi = sizes.clone(), sizes.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.process_group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.process_group)
        if not (torch.equal(lo, sizes) and torch.equal(hi, sizes)):
            raise L.HriemoError(f"Trainer: the ranks hold different models (arena {self.numel} elements / {len(self.slots)} "
                                f"tensors here; min {lo.tolist()}, max {hi.tolist()} over the ranks)")
        dist.broadcast(self.params, src=dist.get_global_rank(self.process_group, 0) if self.process_group is not None else 0,
                       group=self.process_group)

    def _early_layout(self):
        """(names whose gradients are final before encoder layer 0's backward, arena offset where they start) or
        (None, 0) when they do not form the tail of the arena."""
        late = [n for n in self.slots if n.startswith("cross_modal.layers.0.")]
        early = [n for n in self.slots if not n.startswith("cross_modal.layers.0.")]
        if not late or not early:
            return None, 0
        tail_off = min(self.slots[n][0] for n in early)
        if max(self.slots[n][0] + self.slots[n][1] for n in late) > tail_off:
            return None, 0
        return early, tail_off

    def _exchanging(self) -> bool:
        if not (self.distributed and self.overlap):
            return False
        import torch.distributed as dist
        return dist.get_world_size(self.process_group) > 1

    def _exchange_tail_async(self) -> None:
        """All-reduce of the arena's tail, started behind what the current stream holds; on NCCL it runs on the
        process group's own stream and the current stream goes on (it waits in _finish_exchange)."""
        self._early_work = self._all_reduce_mean(self.grads[self._tail_off:], async_op=True, timed=False)

    def _finish_exchange(self) -> None:
        """Layer 0's share, then the join with the tail's all-reduce; the events bracket what the step still waits for."""
        ev = None
        if self.allreduce_events is not None and self.grads.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self._all_reduce_mean(self.grads[:self._tail_off], timed=False)
        if self._early_work is not None:
            self._early_work.wait()
            self._early_work = None
        if ev is not None:
            ev[1].record()
            self.allreduce_events.append(ev)
        self._reduced = True

    def gradient(self, name: str) -> torch.Tensor:
        """View of one parameter's gradient in the arena (as of the last step, after the all-reduce, before the clip)."""
        o, n = self.slots[name]
        return self.grads[o:o + n].view(dict(self.model.named_parameters())[name].shape)

    def _forward_backward(self, h_a, h_t, mask_a, mask_t, labels, scale: float = 1.0, accumulate: bool = False,
                          exchange: bool = False, on_early=None, finish: bool = True) -> dict:
        """Forward with tapes, backward, gradients written (or, with accumulate, added) into the arena, multiplied by
        `scale`.  -> loss / logits / beta / z.
        exchange: data-parallel step with the overlapped exchange -- the gradients that are final before encoder layer 0's
        backward go into the arena at that point and `on_early()` runs (default: start their all-reduce); `finish`: exchange
        layer 0's share at the end and join (a captured step leaves both to the replaying code)."""
        def fill(names, grads):   # device-to-device into the arena: ONE multi-tensor launch instead of a copy per parameter
            dsts = [self.grads[self.slots[n][0]:self.slots[n][0] + self.slots[n][1]] for n in names]
            srcs = [grads[n].reshape(-1) for n in names]
            if accumulate:
                torch._foreach_add_(dsts, srcs, alpha=scale)
            else:
                torch._foreach_copy_(dsts, srcs)
                if scale != 1.0:
                    torch._foreach_mul_(dsts, scale)

        filled = set()
        hook = None
        if exchange:
            if accumulate or scale != 1.0:
                raise L.HriemoError("Trainer: the overlapped exchange is for plain steps (no accumulation)")

            def hook(early):
                if set(early) != set(self._early_names):
                    raise L.HriemoError(f"Trainer: early gradient names do not match: {sorted(set(early) ^ set(self._early_names))[:6]}")
                fill(self._early_names, early)
                filled.update(self._early_names)
                (on_early or self._exchange_tail_async)()
        out = backward.loss_and_gradients(self.model, h_a, h_t, mask_a, mask_t, labels, self.beta_weight, early_hook=hook)
        grads = out.pop("grads")
        if set(grads) != set(self.slots):
            raise L.HriemoError(f"Trainer: gradient names do not match the parameters: {sorted(set(grads) ^ set(self.slots))[:6]}")
        fill([n for n in self.slots if n not in filled], grads)
        if exchange and finish:
            self._finish_exchange()
        return out

    def _graphed_forward_backward(self, h_a, h_t, mask_a, mask_t, labels, exchange: bool = False) -> dict:
        ins = (h_a, h_t, mask_a, mask_t, labels)
        key = tuple(None if x is None else (tuple(x.shape), x.dtype, x.device) for x in ins) + (exchange,)
        if key != self._graph_key:
            self._graph, self._graph_key, self._graph_seen = None, key, 0
        if self._graph is None:
            self._graph_seen += 1
            if self._graph_seen <= 2:                  # eager first: per-device kernel attributes, allocator warm-up
                return self._forward_backward(*ins, exchange=exchange)
            self._static_in = [None if x is None else x.clone() for x in ins]
            invalidate_prepared(self.model)            # so that the weight casts are part of the captured work
            torch.cuda.synchronize()
            if not exchange:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._graph_out = self._forward_backward(*self._static_in)
            else:
                # two graphs sharing one memory pool, split where the tail's all-reduce is launched: [forward, backward
                # down to encoder layer 1, arena fill of the tail] | [layer 0's backward, arena fill of the head]
                import gc
                g_a, g_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                gc.collect()
                torch.cuda.empty_cache()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    g_a.capture_begin()

                    def split():
                        g_a.capture_end()
                        g_b.capture_begin(pool=g_a.pool())
                    self._graph_out = self._forward_backward(*self._static_in, exchange=True, on_early=split, finish=False)
                    g_b.capture_end()
                torch.cuda.current_stream().wait_stream(side)
                self._graph = (g_a, g_b)
        else:
            for dst, src in zip(self._static_in, ins):
                if dst is not None:
                    dst.copy_(src)
        if isinstance(self._graph, tuple):
            self._graph[0].replay()
            self._exchange_tail_async()
            self._graph[1].replay()
            self._finish_exchange()
        else:
            self._graph.replay()
        return dict(self._graph_out)

    def step(self, h_a: torch.Tensor, h_t: torch.Tensor, mask_a, mask_t, labels: torch.Tensor) -> dict:
        dropping = self.p_drop > 0 and self.model.training   # fresh mask keys every step: not replayable
        fb = self._graphed_forward_backward if self.use_graph and not dropping else self._forward_backward
        out = fb(h_a, h_t, mask_a, mask_t, labels, exchange=self._exchanging())
        out.update(self.apply())
        return out

    def apply(self) -> dict:
        """All-reduce (data parallel), global-norm clip and AdamW on what the gradient arena holds; the learning rate is
        `self.lr` at the time of the call (set it per step for a schedule, e.g. the cosine + warm-up of
        train_mosei_fusion_seq_level_decoder.py:574-590).  -> {"grad_norm", "clip"} (device tensors)."""
        if self.distributed and not self._reduced:     # (a step with the overlapped exchange has reduced the arena already)
            self._all_reduce_mean(self.grads)
        self._reduced = False
        norm_clip = ops.grad_norm_clip(self.grads, self.max_norm)          # [total norm, clip coefficient], on the device
        self.step_count += 1
        ops.adamw_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.step_count, lr=self.lr,
                       betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, grad_scale=norm_clip[1:])
        invalidate_prepared(self.model)   # eager callers (model.eval()(...)) re-cast; a captured step re-casts by itself
        return dict(grad_norm=norm_clip[0], clip=norm_clip[1])

    def step_accumulated(self, micro_batches) -> dict:
        """Gradient accumulation as in train_mosei_fusion_seq_level_decoder.py:388-401 (`loss / grad_accum` per micro-batch,
        one clip + optimizer step per `grad_accum` micro-batches).  micro_batches: sequence of (h_a, h_t, mask_a, mask_t,
        labels).  -> {"loss": mean of the micro-batch losses, "grad_norm", "clip"}."""
        micro_batches = list(micro_batches)
        if not micro_batches:
            raise L.HriemoError("Trainer.step_accumulated: no micro-batches")
        k = len(micro_batches)
        loss = None
        for i, mb in enumerate(micro_batches):
            out = self._forward_backward(*mb, scale=1.0 / k, accumulate=i > 0)
            loss = out["loss"] / k if loss is None else loss + out["loss"] / k
        res = self.apply()
        res["loss"] = loss
        return res

    def _all_reduce_mean(self, t: torch.Tensor, async_op: bool = False, timed: bool = True):
        """Mean over the ranks, in place.  async_op (NCCL only; other backends reduce at once): returns the Work whose
        wait() joins the current stream with the collective."""
        import torch.distributed as dist

        world = dist.get_world_size(self.process_group)
        if world == 1 or t.numel() == 0:
            return None
        ev = None
        if timed and self.allreduce_events is not None and t.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        work = None
        if dist.get_backend(self.process_group) == "nccl":
            work = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.process_group, async_op=async_op)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.process_group)
            t.mul_(1.0 / world)
        if ev is not None:
            ev[1].record()
            self.allreduce_events.append(ev)
        return work if async_op else None
