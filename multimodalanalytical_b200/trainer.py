"""Built-in training / prediction loop for the accelerated path.

Replaces what Lightning does around `HFWrapper` in the reference (trainer/trainer.py:9-73: DDP, gradient
accumulation, clip 1.0, optimiser + OneCycleLR stepping) with a fused schedule:
  forward -> backward (gradient buckets all-reduced over NCCL on a side stream as they complete, back to
  front) -> global-norm clip + Adam/AdamW + bf16-mirror refresh + grad zeroing in one kernel pass.
Lightning itself still works with the module (wrapper.py); this loop is what bench.py measures.
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, Optional, List

import torch

from . import ops
from .wrapper import HFWrapper


def one_cycle(step: int, total_steps: int, max_lr: float, pct_start: float = 0.3, div_factor: float = 25.0,
              final_div_factor: float = 1e4, base_momentum: float = 0.85, max_momentum: float = 0.95):
    """(lr, beta1) of torch.optim.lr_scheduler.OneCycleLR (defaults, cos annealing, cycle_momentum=True) after
    `step` scheduler steps - what the reference gets from OneCycleLR(optim, lr, total_steps) (wrapper.py:341)."""
    initial_lr = max_lr / div_factor
    min_lr = initial_lr / final_div_factor
    p1_end = float(pct_start * total_steps) - 1
    p2_end = total_steps - 1

    def cos(start, end, pct):
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1)

    if step <= p1_end:
        pct = step / p1_end if p1_end > 0 else 1.0
        return cos(initial_lr, max_lr, pct), cos(max_momentum, base_momentum, pct)
    pct = (step - p1_end) / (p2_end - p1_end) if p2_end > p1_end else 1.0
    pct = min(pct, 1.0)
    return cos(max_lr, min_lr, pct), cos(base_momentum, max_momentum, pct)


DDP_DEFAULT = "p2p"  # MMA_DDP=nccl: bucketed NCCL all-reduce overlapped with backward + replicated Adam


class GradBucketer:
    """Back-to-front bucketed all-reduce of a flat gradient buffer.

    The engine's backward reports `on_ready(off)`: every gradient at flat offset >= off is final.  Full buckets are
    all-reduced (SUM) right away on a side stream (NCCL) so the transfer overlaps the rest of backward; `finish()`
    joins the side stream.  On CPU tensors (gloo; used by the tests) the reduction is synchronous."""

    def __init__(self, flat_grad: torch.Tensor, bucket_elems: int, process_group=None):
        self.g, self.bucket, self.pg = flat_grad, int(bucket_elems), process_group
        self.comm_stream = torch.cuda.Stream() if flat_grad.is_cuda else None
        self.hi = flat_grad.numel()
        self.launched = []

    def reset(self):
        self.hi = self.g.numel()
        self.launched = []

    def on_ready(self, off: int, also_wait=None):
        """`also_wait`: an event of another producer stream (the engine's weight-gradient side stream) that must
        have completed before the reduced range is read."""
        while self.hi - off >= self.bucket or (off == 0 and self.hi > 0):
            lo = max(off, self.hi - self.bucket)
            self._allreduce(lo, self.hi, also_wait)
            self.hi = lo

    def _allreduce(self, lo, hi, also_wait=None):
        g = self.g[lo:hi]
        self.launched.append((lo, hi))
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.comm_stream.wait_event(ev)
            if also_wait is not None:
                self.comm_stream.wait_event(also_wait)
            with torch.cuda.stream(self.comm_stream):
                torch.distributed.all_reduce(g, group=self.pg)
        else:
            torch.distributed.all_reduce(g, group=self.pg)

    def finish(self, also_wait=None):
        if self.hi > 0:
            self.on_ready(0, also_wait)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


class PeerShardedStep:
    """Gradient reduction + optimiser step over NVLink peer memory (csrc/ddp_p2p.cu): barrier -> each rank sums ITS shard
    of the gradients from all peers -> barrier -> Adam on the shard, bf16 weights stored into every peer's mirror ->
    barrier.  The gradient buffer, the bf16 mirror, the partial-norm slot and the barrier flags live in symmetric memory
    (`torch.distributed._symmetric_memory`: torch allocates and exchanges the handles - plumbing; every byte on the data
    path is moved by this repo's kernels).  Master weights and Adam moments are sharded by rank afterwards
    (`ParamStore.gather_master` reassembles them for checkpoints)."""

    def __init__(self, store, process_group=None):
        import torch.distributed._symmetric_memory as symm

        pg = process_group if process_group is not None else torch.distributed.group.WORLD
        self.pg = pg
        self.world, self.rank = torch.distributed.get_world_size(pg), torch.distributed.get_rank(pg)
        dev, n = store.device, store.numel
        self.g = symm.empty(n, dtype=torch.float32, device=dev)
        self.pb = symm.empty(n, dtype=torch.bfloat16, device=dev)
        self.flags = symm.empty(16, dtype=torch.int32, device=dev).zero_()
        self.sumsq = symm.empty(4, dtype=torch.float32, device=dev).zero_()
        name = pg.group_name
        self._remote = []  # keeps the peers' mapped views alive

        self.mc = {}  # NVLS multicast address of a symmetric tensor (0: not available)
        use_mc = os.environ.get("MMA_DDP_MULTICAST", "1") != "0"

        def peers(t):
            hdl = symm.rendezvous(t, name)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if use_mc else 0
            # same offset inside the multicast mapping as inside this rank's own allocation
            self.mc[t.data_ptr()] = mc + (t.data_ptr() - int(hdl.buffer_ptrs[self.rank])) if mc else 0
            out = []
            for r in range(self.world):
                if r == self.rank:
                    out.append(t.data_ptr())
                else:  # the peer's copy of THIS tensor (its offset inside the symmetric allocation included)
                    rt = hdl.get_remote_tensor(r, tuple(t.shape), t.dtype)
                    self._remote.append(rt)
                    out.append(rt.data_ptr())
            return out

        self.peer_g, self.peer_pb = peers(self.g), peers(self.pb)
        self.peer_flags, self.peer_sumsq = peers(self.flags), peers(self.sumsq)
        # address check before any kernel trusts the table: every rank publishes its id, every rank reads every peer's
        self.sumsq.fill_(float(self.rank + 1))
        torch.cuda.synchronize()
        torch.distributed.barrier(group=pg)
        for rt_rank, rt in zip([r for r in range(self.world) if r != self.rank], self._remote[-(self.world - 1):]):
            got = float(rt[0].item())
            if got != float(rt_rank + 1):
                raise RuntimeError(f"peer-memory address table is wrong: rank {rt_rank} slot reads {got}")
        torch.distributed.barrier(group=pg)
        self.sumsq.zero_()
        store.rebind(g=self.g, pb=self.pb)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ws = torch.zeros(1024, dtype=torch.float32, device=dev)
        per = -(-n // self.world)
        per = (per + 127) // 128 * 128
        self.lo, self.hi = min(n, self.rank * per), min(n, (self.rank + 1) * per)
        store.shard = (self.lo, self.hi)
        torch.cuda.synchronize()
        torch.distributed.barrier(group=pg)  # every rank's flags are zeroed before the first device-side barrier

    def barrier(self):
        ops.p2p_barrier(self.peer_flags, self.epoch, self.world, self.rank)

    def step(self, store, hyper, decoupled: bool):
        self.barrier()
        ops.p2p_reduce_shard(self.peer_g, self.world, self.rank, self.lo, self.hi, self.ws, self.sumsq,
                             mc_g=self.mc.get(self.g.data_ptr(), 0))
        self.barrier()
        ops.p2p_adam_shard(store.p, store.g, store.m, store.v, self.peer_pb, self.peer_sumsq, self.world, self.rank,
                           self.lo, self.hi, hyper, decoupled=decoupled, mc_pb=self.mc.get(self.pb.data_ptr(), 0))
        self.barrier()


class FusedTrainer:
    BUCKET_ELEMS = 8 * 1024 * 1024  # 32 MB of fp32 gradients per all-reduce
    HYPER_SLOTS = 64
    MAX_GRAPHS = 12  # distinct input-shape signatures kept as captured steps (each owns a step's worth of workspaces)

    def __init__(self, module: HFWrapper, clip_grad: float = 1.0, acc_batches: int = 1, process_group=None,
                 eps: float = 1e-8, use_graph: bool = True):
        self.m = module
        self.use_graph = use_graph
        self._graphs: Dict[Any, Any] = {}
        self.eng, self.ps = module.engine, module.store
        self.clip_grad, self.acc, self.eps = clip_grad, max(1, acc_batches), eps
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if self._dist() else 1
        self.opt_step = 0
        self.micro = 0
        dev = self.ps.device
        self.ps.ensure_optimizer_state()
        # ring of pinned staging rows: the async H2D of step i must not be overwritten by a later step's host writes.
        # Graph replay lets the host run far ahead of the GPU, so every slot carries an event recorded after its copy
        # and is only rewritten once that event has completed.
        self.hyper_ring = torch.zeros(self.HYPER_SLOTS, 9, dtype=torch.float32)
        self._hyper_ev: List[Any] = [None] * self.HYPER_SLOTS
        if dev.type == "cuda":
            self.hyper_ring = self.hyper_ring.pin_memory()
        self.hyper = torch.zeros(9, dtype=torch.float32, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.norm_ws = torch.zeros(1024, dtype=torch.float32, device=dev)
        # multi-GPU gradient exchange: "p2p" = reduce-scatter + sharded Adam + bf16 all-gather by this repo's kernels over
        # NVLink peer memory (bf16 engines; needs clipping semantics of the global norm: kept), "nccl" = bucketed NCCL
        # all-reduce overlapped with backward + replicated Adam
        self.ddp_mode = os.environ.get("MMA_DDP", DDP_DEFAULT)
        if self.eng.precision != "bf16" or dev.type != "cuda":
            self.ddp_mode = "nccl"
        self.peer = None
        if self.world > 1 and self.ddp_mode == "p2p":
            # every rank must take the same path: agree on whether the symmetric allocation + address exchange worked
            ok = torch.ones(1, device=dev)
            try:
                self.peer = PeerShardedStep(self.ps, process_group)
            except Exception as exc:  # noqa: BLE001  (no peer access / no symmetric-memory support on this system)
                import warnings

                warnings.warn(f"peer-memory optimiser step unavailable ({exc!r}); using the NCCL all-reduce path")
                ok.zero_()
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=process_group)
            if float(ok) == 0.0:
                if self.peer is not None:
                    self.ps.shard = None
                self.peer, self.ddp_mode = None, "nccl"
        self.bucketer = GradBucketer(self.ps.g, self.BUCKET_ELEMS, process_group) if self.world > 1 and self.peer is None else None
        # multi-GPU: the bucketed NCCL all-reduces (forked onto the side stream) are captured into the step's CUDA
        # graph together with the kernels; MMA_DDP_GRAPH=0 falls back to eager launches
        self.graph_ddp = os.environ.get("MMA_DDP_GRAPH", "1") != "0"
        self._sync_now = False
        self._flush_next = False  # fit(): the last batch of an epoch steps the optimiser even inside a partial window
        self.eng.grad_ready_hook = self._on_grads_ready
        self.eng.on_release.append(self._graphs.clear)

    def _dist(self):
        return torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(self.pg) > 1

    # gradient buckets complete back to front; reduce each as soon as it is final
    def _on_grads_ready(self, off: int):
        if self._sync_now and self.bucketer is not None:
            self.bucketer.on_ready(off, also_wait=self.eng.last_wgrad_event)

    def _prepare(self, batch):
        """Collator dict -> device tensors in the engine's layout (what HFWrapper.forward does, wrapper.py:356-389)."""
        m = self.m
        input_ids, attention_mask = m._relayout(batch, training=True)
        dec_in = m._to_dev(batch["decoder_input"][m.target_modality]).transpose(1, 0).contiguous()
        dec_mask = (~m._to_dev(batch["decoder_pad_mask"])).T.to(torch.uint8).contiguous()
        labels = m._to_dev(batch["target"]).T.contiguous().clone()
        labels[labels == m.target_tokenizer.pad_token_id] = -100
        if m.engine.cfg.align_config and "encoder_alignment_input" in batch:
            return (input_ids, attention_mask, dec_in, dec_mask, labels,
                    m._to_dev(batch["encoder_alignment_input"]).float().contiguous())
        return input_ids, attention_mask, dec_in, dec_mask, labels

    @staticmethod
    def _flat(inputs):
        ids, am, di, dm, lb = inputs[:5]
        out = []
        for k, v in ids.items():
            if isinstance(v, dict):
                out += [(f"{k}.{kk}", t) for kk, t in v.items()]
            else:
                out.append((k, v))
        out += [("enc_mask", am), ("dec_in", di), ("dec_mask", dm), ("labels", lb)]
        if len(inputs) > 5:
            out.append(("align_target", inputs[5]))
        return out

    def _step_body(self, inputs):
        ids, am, di, dm, lb = inputs[:5]
        self.eng.next_seed()
        out = self.eng.forward(ids, am, di, dm, labels=lb, train=True, align_target=inputs[5] if len(inputs) > 5 else None)
        self.eng.backward(gscale=1.0)
        return out["loss"]

    def train_step(self, batch: Dict[str, Any], batch_idx: int = 0):
        """One micro-batch: forward + backward (+ optimiser step every `acc_batches`).  Returns the loss scalar
        (device tensor; no host sync).  With `use_graph` the whole step (seed advance, forward, backward, clip,
        Adam) is captured once per input-shape signature and replayed."""
        m = self.m
        m.train()
        if self.ps.g_dirty:  # a backward through the autograd path (HFWrapper.forward + loss.backward()) left gradients
            self.ps.g.zero_()
            self.ps.g_dirty = False
        self.micro += 1
        self._sync_now = (self.micro % self.acc) == 0 or self._flush_next
        self._flush_next = False
        if self._sync_now:
            self.micro = 0
        if self.bucketer is not None:
            self.bucketer.reset()
        # a tuple is a batch already in the engine's layout (pipeline.DeviceDataset.collate)
        inputs = batch if isinstance(batch, tuple) else self._prepare(batch)
        if self.use_graph and self.acc == 1 and (self.bucketer is None or self.graph_ddp):
            return self._graphed_step(inputs)
        loss = self._step_body(inputs)
        if self._sync_now:
            self._write_hyper()
            self._optimizer_kernels()
            self.opt_step += 1
        return loss

    def _graphed_step(self, inputs):
        flat = self._flat(inputs)
        key = tuple((n, tuple(t.shape), t.dtype) for n, t in flat)
        ent = self._graphs.get(key)
        if ent is None:
            # first sight of this shape: run it eagerly (allocates workspaces, builds tensor maps) on PRIVATE copies of
            # the inputs, which become the static buffers of the graph captured on the next step of this shape (the
            # caller's tensors are never adopted or written to)
            if len(self._graphs) >= self.MAX_GRAPHS:
                # captured graphs hold raw pointers into the engine's shape-keyed workspaces, so eviction is
                # all-or-nothing: drop every graph, then every workspace
                torch.cuda.synchronize()
                self.eng.release_buffers()
            static = self._clone_inputs(inputs)
            loss = self._step_body(static)
            self._write_hyper()
            self._optimizer_kernels()
            self.opt_step += 1
            self._graphs[key] = {"static": static, "graph": None, "loss": None}
            return loss.clone()
        for (_, dst), (_, src) in zip(self._flat(ent["static"]), flat):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._write_hyper()
        if ent["graph"] is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(g):
                ent["loss"] = self._step_body(ent["static"])
                self._optimizer_kernels()
            ent["graph"], ent["launches"] = g, ops.LAUNCHES - n0
            ops.LAUNCHES = n0
        ent["graph"].replay()
        ops.LAUNCHES += ent["launches"]  # kernels executed by the replay
        self.opt_step += 1
        return ent["loss"].clone()  # the graph's loss buffer is overwritten by the next replay

    @staticmethod
    def _clone_inputs(inputs):
        def cl(v):
            if isinstance(v, dict):
                return {k: cl(t) for k, t in v.items()}
            return v.clone()

        return tuple(cl(v) for v in inputs)

    def _write_hyper(self):
        """Stage this optimiser step's scalars (OneCycle lr / beta1, bias corrections, clip, grad scale) on the device."""
        m = self.m
        lr, beta1 = one_cycle(self.opt_step, m.num_steps, m.lr)
        if m.num_steps <= 0:
            lr, beta1 = m.lr, m.adam_beta1
        t = self.opt_step + 1
        slot = self.opt_step % self.HYPER_SLOTS
        if self._hyper_ev[slot] is not None:
            self._hyper_ev[slot].synchronize()  # the copy that last read this slot has executed
        h = self.hyper_ring[slot]
        h[0], h[1], h[2], h[3], h[4] = lr, beta1, m.adam_beta2, self.eps, m.weight_decay
        h[5], h[6] = 1.0 - beta1 ** t, 1.0 - m.adam_beta2 ** t
        h[7] = self.clip_grad if self.clip_grad else 0.0
        h[8] = 1.0 / (self.world * self.acc)
        self.hyper.copy_(h, non_blocking=True)
        if self.hyper.is_cuda:
            ev = self._hyper_ev[slot] or torch.cuda.Event()
            ev.record()
            self._hyper_ev[slot] = ev

    def _optimizer_kernels(self):
        ps = self.ps
        if self.peer is not None:
            self.peer.step(ps, self.hyper, decoupled=(self.m.optimiser == "adamw"))
            ps.bf16_dirty = False
            return
        if self.bucketer is not None:
            self.bucketer.finish(also_wait=self.eng.last_wgrad_event)
        norm = None
        if self.clip_grad:
            ops.grad_norm(ps.g, self.norm_ws, self.norm)
            norm = self.norm
        ops.adam_step(ps.p, ps.g, ps.m, ps.v, ps.pb if self.eng.precision == "bf16" else None, self.hyper, norm=norm,
                      decoupled=(self.m.optimiser == "adamw"), zero_grad=True)
        ps.bf16_dirty = False

    def optimizer_step(self):
        self._write_hyper()
        self._optimizer_kernels()
        self.opt_step += 1

    def fit(self, batches, epochs: int = 1, log_every: int = 10, log=print):
        step = 0
        for ep in range(epochs):
            if hasattr(batches, "set_epoch"):  # pipeline.DeviceLoader / samplers: new shuffle per epoch
                batches.set_epoch(ep)
            # Lightning steps the optimiser on the last batch of every epoch even when the accumulation window is not
            # full (ceil(batches / acc) optimiser steps per epoch = what calculate_training_steps gives OneCycleLR):
            # look one batch ahead so the last one is known
            it = iter(batches)
            nxt = next(it, None)
            i = 0
            while nxt is not None:
                batch, nxt = nxt, next(it, None)
                self._flush_next = nxt is None
                loss = self.train_step(batch, i)
                if log and step % log_every == 0:
                    log(f"epoch {ep} step {step} train_loss {float(loss):.4f}")
                step += 1
                i += 1
        return step


@torch.no_grad()
def predict(module: HFWrapper, batches, n_beams: Optional[int] = None):
    """trainer.predict(model, datamodule) of the reference CLI (cli/training.py:190-206), minus Lightning."""
    outs = []
    if n_beams is not None:
        module.n_beams = n_beams
    for i, batch in enumerate(batches):
        outs.append(module.predict_step(batch, i))
    return outs


def calculate_training_steps(n_train: int, batch_size: int, acc_batches: int = 1, epochs: int = 1, world: int = 1,
                             reference_compat: bool = True) -> int:
    """Optimiser steps of a run = OneCycleLR's `total_steps` (analytical_fm/utils.py:155-172).  The reference divides
    by a hard-coded GPU count of 1, so under DDP its schedule is `world` times longer than the steps actually taken
    (SURVEY App. A #10); `reference_compat=True` keeps that, False divides by the real world size."""
    gpus = 1 if reference_compat else max(1, world)
    batches_per_gpu = math.ceil((n_train / batch_size) / float(gpus))
    return math.ceil(batches_per_gpu / acc_batches) * epochs


def shard_indices(n: int, rank: int, world: int, contiguous: bool = False):
    """Sample indices of `rank` for sharded inference (independent spectra per GPU, no data-path
    collective).  Strided by default (rank r takes r, r + world, ...: DistributedSampler order without padding, so
    no sample is decoded twice); `contiguous=True` gives rank r the r-th block."""
    if not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    if contiguous:
        per = -(-n // world)
        return list(range(min(n, rank * per), min(n, (rank + 1) * per)))
    return list(range(rank, n, world))


def gather_outputs(local: List[Any], indices: List[int], n: int, process_group=None) -> Optional[List[Any]]:
    """Reassemble per-sample outputs of a sharded predict in dataset order.  Host objects only (decoded strings,
    losses), exchanged once at the end with `all_gather_object` - nothing on the decode path waits on another rank
    (the reference writes one pickle per rank instead, cli/training.py:230-240).  Every rank returns the full list."""
    if len(local) != len(indices):
        raise ValueError("one output per local index expected")
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        parts = [(indices, local)]
    else:
        world = torch.distributed.get_world_size(process_group)
        parts = [None] * world
        torch.distributed.all_gather_object(parts, (indices, local), group=process_group)
    out: List[Any] = [None] * n
    seen = 0
    for idx, vals in parts:
        for i, v in zip(idx, vals):
            if out[i] is not None:
                raise ValueError(f"sample {i} was produced by two ranks")
            out[i] = v
            seen += 1
    if seen != n:
        raise ValueError(f"{n - seen} samples were not produced by any rank")
    return out


@torch.no_grad()
def predict_sharded(module: HFWrapper, dataset, batch_size: int, rank: int = 0, world: int = 1,
                    n_beams: Optional[int] = None, process_group=None) -> List[List[str]]:
    """Beam-search prediction of a `pipeline.DeviceDataset` split over `world` GPUs: each rank decodes its own
    samples (wire batches assembled on its device), then the decoded hypotheses are exchanged once.  Returns, on every
    rank, `n_beams` strings per sample in dataset order."""
    K = n_beams if n_beams is not None else module.n_beams
    mine = shard_indices(len(dataset), rank, world)
    local: List[List[str]] = []
    module.eval()
    for lo in range(0, len(mine), batch_size):
        batch = dataset.wire_batch(mine[lo: lo + batch_size])
        seqs = module.generate(batch, n_beams=K)
        dec = module.target_tokenizer.batch_decode(seqs, skip_special_tokens=True)
        local += [dec[i * K: (i + 1) * K] for i in range(len(dec) // K)]
    return gather_outputs(local, mine, len(dataset), process_group)


@torch.no_grad()
def validate(module: HFWrapper, batches, limit_batches: Optional[int] = None) -> Dict[str, float]:
    """One validation epoch: what Lightning runs around `validation_step` / `on_validation_epoch_end`
    (wrapper.py:491-530): per batch the teacher-forced loss, token accuracy, a KV-cached greedy decode and its Top-1
    molecular accuracy; returns the epoch means.  Host syncs happen once per batch (string comparison of the decoded
    SMILES), never per decode step."""
    was_training = module.training
    sums: Dict[str, float] = {}
    n = 0
    for i, batch in enumerate(batches):
        if limit_batches is not None and i >= limit_batches:
            break
        out = module.validation_step(batch, i)
        for k, v in out.items():
            if v is not None:
                sums[k] = sums.get(k, 0.0) + float(v)
        n += 1
    module.on_validation_epoch_end()
    if was_training:
        module.train()
    return {k: v / max(n, 1) for k, v in sums.items()}
