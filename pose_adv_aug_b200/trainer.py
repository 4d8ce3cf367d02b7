"""The whole reference train step (stack-hg.py:153-165: forward -> sum over stacks of MSE ->
zero_grad / backward / RMSprop.step) as one planned launch sequence, optionally captured in a
single CUDA graph: weight repack, every conv / BN / pool / up-add kernel, the fused MSE
forward+backward, the backward pass, the (optional) NCCL all-reduce of the flat gradient buffer
and the flat RMSprop update."""
import os

import torch

from ._lib import get_lib, HGKError
from .engine import Plan
from .models import asn_stacked_hg as M
from . import dist as hdist


class _HyperGroup(dict):
    """param_groups[0] of the trainer: a dict whose 'lr' / 'alpha' / 'eps' writes go straight to the 4-float device array the
    update kernel reads (no graph re-capture, no host synchronisation beyond the small copy)."""
    _SLOT = {"lr": 0, "alpha": 1, "eps": 2}

    def __init__(self, owner, lr, alpha, eps):
        dict.__init__(self, lr=lr, alpha=alpha, eps=eps)
        self._owner = owner

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, value)
        slot = self._SLOT.get(key)
        if slot is not None:
            self._owner.hyper[slot:slot + 1].fill_(float(value))

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v


def bucket_splits(offsets, numel, nb):
    """Element offsets cutting a flat gradient buffer into (at most) `nb` buckets of about equal size, each cut on a
    parameter-slot boundary (so that every kernel's gradient write falls into exactly one bucket)."""
    target = numel / float(nb)
    splits = []
    for o in offsets[1:]:
        if len(splits) < nb - 1 and o >= target * (len(splits) + 1):
            splits.append(o)
    return splits


def param_completion(plan, store):
    """Index in plan.bwd of the last launch that contributes to each parameter's gradient (-1: none), read off the emitted
    backward list before the weight-gradient scratch references are resolved."""
    import bisect
    from .engine import _WRITES, _WgRef
    g0, g1 = store.grad.data_ptr(), store.grad.data_ptr() + 4 * store.numel
    offs = store.offsets
    wg_param = {}
    for (w, off) in plan.wg_entries:
        a = plan.param_grad_ptr(w)
        if g0 <= a < g1:
            wg_param[off] = bisect.bisect_right(offs, (a - g0) // 4) - 1
    pos = [-1] * len(offs)
    for i, rec in enumerate(plan.bwd):
        wpos = _WRITES.get(rec[2])
        for j, a in enumerate(rec[1]):
            if isinstance(a, _WgRef):
                if a.off in wg_param:
                    pos[wg_param[a.off]] = i
            elif isinstance(a, int) and g0 <= a < g1 and (wpos is None or j in wpos):
                pos[bisect.bisect_right(offs, (a - g0) // 4) - 1] = i
    return pos


def plan_bucket_splits(plan, store, nb):
    """Cut the flat gradient buffer into `nb` contiguous buckets so that as many bytes as possible are final EARLY in the
    backward pass: minimise sum over buckets of (bucket elements x position of the bucket's last writer) by dynamic
    programming over the parameter slots.  Parameters are laid out in registration order and the backward pass finishes
    them roughly back to front (second hourglass, first hourglass' up path, its skips, its down path, the stem), so the
    buckets follow those groups and the all-reduce exposed behind the last kernel is the small stem bucket."""
    pos = param_completion(plan, store)
    P = len(pos)
    nb = max(1, min(nb, P))
    edges = list(store.offsets) + [store.numel]
    n_bwd = float(max(len(plan.bwd), 1))
    c = [(p + 1) / n_bwd for p in pos]
    INF = float("inf")
    best = [[INF] * (P + 1) for _ in range(nb + 1)]
    arg = [[0] * (P + 1) for _ in range(nb + 1)]
    best[0][0] = 0.0
    for k in range(1, nb + 1):
        for j in range(1, P + 1):
            cmax = 0.0
            for i in range(j - 1, -1, -1):           # bucket = parameters i .. j-1
                if c[i] > cmax:
                    cmax = c[i]
                if best[k - 1][i] < INF:
                    v = best[k - 1][i] + (edges[j] - edges[i]) * cmax
                    if v < best[k][j]:
                        best[k][j], arg[k][j] = v, i
    k = min(range(1, nb + 1), key=lambda q: best[q][P])
    cuts, j = [], P
    while k > 0:
        i = arg[k][j]
        if i > 0:
            cuts.append(edges[i])
        j, k = i, k - 1
    return sorted(cuts)


def bucket_last_writers(launches, rw_override, grad_ptr, numel, splits):
    """For each gradient bucket the index of the last launch (list order = a topological order) that writes into it."""
    import bisect
    from .engine import _WRITES
    g0, g1 = grad_ptr, grad_ptr + 4 * numel
    last = [-1] * (len(splits) + 1)
    for i, rec in enumerate(launches):
        ov = rw_override.get(id(rec))
        if ov is not None:
            wr = ov[1]
        else:
            wpos = _WRITES.get(rec[2])
            wr = [a for j, a in enumerate(rec[1]) if isinstance(a, int) and (wpos is None or j in wpos)]
        for a in wr:
            if g0 <= a < g1:
                last[bisect.bisect_right(splits, (a - g0) // 4)] = i
    return last


class HourglassTrainer(object):
    def __init__(self, net, batch, res, lr=2.5e-4, alpha=0.99, eps=1e-8, device=None, use_graph=True,
                 distributed=None, n_streams=10, n_low=3, ar_buckets=None, fake_collective=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.net, self.N, self.R, self.device = net, batch, res, torch.device(device)
        self.lib = get_lib()
        self.world = hdist.world_size() if distributed is None else (hdist.world_size() if distributed else 1)
        net.to(self.device)
        net.train()
        self.store = M._ensure_store(net, self.device)
        if self.world > 1:
            hdist.broadcast_flat_params(self.store.flat)
        self.square_avg = torch.zeros_like(self.store.flat)
        # (lr, alpha, eps, 1/world) live in DEVICE memory and are read by the update kernel at execution time: the captured
        # CUDA graph follows later changes (the reference's adjust_lr, stack-hg.py:106 / utils/util.py:105) without re-capture
        self.hyper = torch.tensor([lr, alpha, eps, 1.0 / self.world], device=self.device, dtype=torch.float32)
        self.param_groups = [_HyperGroup(self, lr, alpha, eps)]      # torch.optim-style access: param_groups[0]['lr'] = ...
        K = net.num_classes
        self.x = torch.zeros(batch, 3, res, res, device=self.device)
        self.t = torch.zeros(batch, K, res // 4, res // 4, device=self.device)
        self.loss_acc = torch.zeros(1, device=self.device, dtype=torch.float64)
        self.loss = torch.zeros(1, device=self.device, dtype=torch.float32)
        plan = Plan([self.store], self.device, True, True, conv_path=M.CONV_PATH, precise_grads=M.PRECISE_GRADS)
        img = plan.input_image(batch, res, res)
        tgt = plan.target_nchw(batch, K, res // 4, res // 4)
        outs, _ = net._build(plan, img)
        for o in outs:
            plan.mse_loss(o, tgt, self.loss_acc)
            plan.output_nchw(o, no_grad=True)
        # gradient buckets of the in-graph all-reduce: contiguous ranges of the flat gradient buffer cut at parameter-slot
        # boundaries.  Parameters are laid out in registration order and the backward pass completes them back to front, so
        # the LAST bucket is final first and its all-reduce runs under the rest of the backward.
        nb = int(os.environ.get("HGK_AR_BUCKETS", "5")) if ar_buckets is None else int(ar_buckets)
        self.ar_in_graph = (self.world > 1 or fake_collective is not None) and use_graph and n_streams > 1 and nb > 0 \
            and os.environ.get("HGK_AR_INGRAPH", "1") == "1"
        self.fake_collective = fake_collective
        plan.finish(grad_splits=(lambda pl: plan_bucket_splits(pl, self.store, nb)) if self.ar_in_graph else None)
        self.grad_splits = plan.grad_splits
        self.plan = plan
        self.graph = None
        self.graph_update = None
        self._sched = None
        self.n_streams = n_streams
        self.n_low = n_low          # of those, low-priority streams reserved for the weight-gradient kernels
        self.use_graph = use_graph
        self.steps = 0

    # number of libhgk kernel launches per step (for bench.py's gpu_launches)
    @property
    def launches_per_step(self):
        return len(self.plan.head_launches()) + len(self.plan.fwd) + len(self.plan.bwd) + 2

    def bucket_ranges(self):
        """[(lo, hi)) element ranges of the gradient buckets (one range when the all-reduce is not bucketed)."""
        edges = [0] + list(self.grad_splits or []) + [self.store.numel]
        return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]

    def _bucket_last_writers(self, launches):
        return bucket_last_writers(launches, self.plan.rw_override, self.store.grad.data_ptr(), self.store.numel,
                                   self.grad_splits or [])

    def _allreduce_bucket(self, lo, hi):
        g = self.store.grad[lo:hi]
        if os.environ.get("HGK_AR_SKIP", "0") == "1":
            return          # TIMING DIAGNOSTIC ONLY (wrong gradients): no collective, so every rank runs at its own pace
        if self.fake_collective is not None:
            self.fake_collective(g)
        else:
            hdist.allreduce_flat_grads(g)

    def _body_grads(self):
        """zero grads/loss -> forward -> fused MSE -> backward (local gradients in the flat buffer)."""
        st, plan = self.store, self.plan
        self.loss_acc.zero_()
        st.grad.zero_()
        plan.run_forward([self.x, self.t])
        plan.run_backward([self.x, self.t], [None] * len(plan.outputs))

    def _body_grads_multistream(self, n_streams=4):
        """Same work as _body_grads, but the static launch list is spread over several streams according to
        its data dependencies (engine.schedule_streams).  Only used under CUDA-graph capture, where the
        fork/join events become graph edges."""
        from .engine import schedule_streams
        st, plan = self.store, self.plan
        main = torch.cuda.current_stream(self.device)
        self.loss_acc.zero_()
        st.grad.zero_()
        for (key, t), x in zip(plan.inputs, [self.x, self.t]):
            plan.patch(key, x.data_ptr())
        if plan.stat_f_used:
            plan.stat_f[:plan.stat_f_used].zero_()
        if plan.stat_b_used:
            plan.stat_b[:plan.stat_b_used].zero_()
        if plan.wg_buf is not None:
            plan.wg_buf.zero_()
        for op in plan.outputs:
            if not op.no_grad and op.gsrc is not None:
                op.gsrc.zero_()
                plan.patch("gout%d" % op.index, op.gsrc.data_ptr())
        for a in plan.aux_zero:
            a.zero_()
        early_stem = os.environ.get("HGK_EARLY_STEM", "0") == "1"      # measured neutral (10.10 vs 10.09 ms): the persistent stem grid leaves no room next to it
        launches = plan.step_launches() if early_stem else plan.head_launches() + plan.fwd + plan.bwd
        n_low = min(self.n_low, n_streams - 1)
        if self._sched is None:
            low_on = os.environ.get("HGK_LOW_SKIPS", "0") == "1" or M.DEFER_SKIPS
            low_ids = plan.low_recs if low_on else ()
            self._sched = schedule_streams(launches, n_streams, n_low=n_low, low_ids=low_ids,
                                           after=plan.after if M.DEFER_SKIPS else None, rw_override=plan.rw_override,
                                           no_pack_ids=plan.no_pack_dep,
                                           low_names=("conv_wgrad_tc_nhwc", "conv_wgrad_nhwc", "stem_conv7_wgrad", "stem_conv7_wgrad_bnapply"))
        stream_of, cross = self._sched
        ar_after = {}
        if self.ar_in_graph:
            for b, i_last in enumerate(self._bucket_last_writers(launches)):
                ar_after.setdefault(max(i_last, 0), []).append(b)
            ranges = self.bucket_ranges()
            comm = torch.cuda.Stream(self.device, priority=-1)
        # with a low-priority pool the other side streams are high priority (-1); the capture stream keeps priority 0
        side = [torch.cuda.Stream(self.device, priority=(0 if (n_low and i >= n_streams - 1 - n_low) else (-1 if n_low else 0)))
                for i in range(n_streams - 1)]
        streams = [main] + side
        fork = torch.cuda.Event()
        fork.record(main)
        for s_ in side:
            s_.wait_event(fork)
        events = {}
        users = set(d for c in cross for d in c)
        last_on = [False] * n_streams          # the last operation on stream k was a libhgk kernel launch
        arm = self.lib.pdl_arm
        for i, (fn, args, name) in enumerate(launches):
            k = stream_of[i]
            sk = streams[k]
            for d in cross[i]:
                sk.wait_event(events[d])
            # programmatic dependent launch (hgk.h): only directly behind another launch of the same stream; a launch that
            # also waits for events of other streams keeps plain (full) dependencies
            arm(1 if (last_on[k] and not cross[i]) else 0)
            last_on[k] = True
            rc = fn(*args, sk.cuda_stream)
            if rc != 0:
                raise HGKError("%s failed (%d): %s" % (name, rc, self.lib.last_error()))
            if i in users:
                ev = torch.cuda.Event()
                ev.record(sk)
                events[i] = ev
            for b in ar_after.get(i, ()):
                # bucket b of the flat gradient buffer is final once everything issued so far has run (list order is a
                # topological order): the communication stream joins every compute stream here and all-reduces the bucket
                # while the streams go on with the rest of the backward pass (captured: NCCL kernel nodes in the step graph)
                for s_ in streams:
                    ev = torch.cuda.Event()
                    ev.record(s_)
                    comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    self._allreduce_bucket(*ranges[b])
        if self.ar_in_graph:
            ev = torch.cuda.Event()
            ev.record(comm)
            main.wait_event(ev)
        for s_ in side:                      # join
            ev = torch.cuda.Event()
            ev.record(s_)
            main.wait_event(ev)
        for f in plan.nbt_flat:
            f.add_(1)
        for b in plan.nbt_bufs:
            b.add_(1)
        plan.fwd_count += 1

    def _body_update(self):
        """flat RMSprop (1/world folded in) + loss accumulator -> fp32 scalar."""
        st = self.store
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.lib.check(self.lib.rmsprop_flat_dev(st.flat.data_ptr(), st.grad.data_ptr(), self.square_avg.data_ptr(),
                                                 st.numel, self.hyper.data_ptr(), stream), "hgk_rmsprop_flat_dev")
        self.lib.check(self.lib.f64_to_f32(self.loss_acc.data_ptr(), self.loss.data_ptr(), 1, 1.0, stream),
                       "hgk_f64_to_f32")

    def _step_body(self):
        self._body_grads()
        if self.world > 1:
            hdist.allreduce_flat_grads(self.store.grad)      # the ONE collective of the step (NCCL, NVLink)
        self._body_update()

    def step_resident(self):
        """One train step on the batch currently held in the static device buffers self.x / self.t."""
        if not self.store.valid():
            raise HGKError("parameters were re-allocated after the trainer was built")
        if self.use_graph:
            if self.graph is None:
                self._capture()
            if self.world > 1 and not self.ar_in_graph:
                # HGK_AR_INGRAPH=0: two graphs around ONE eager NCCL call over the whole flat buffer
                self.graph.replay()
                hdist.allreduce_flat_grads(self.store.grad)
                self.graph_update.replay()
            else:
                self.graph.replay()
        else:
            self._step_body()
        self.steps += 1
        return self.loss

    def _capture(self):
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            saved = (self.store.flat.clone(), self.square_avg.clone(), self.store.fbuf_flat.clone(),
                     self.store.ibuf_flat.clone())
            self._step_body()            # warm-up outside capture (lazy init, NCCL communicator)
            if self.ar_in_graph and self.world > 1:
                scratch = torch.zeros_like(self.store.grad)
                for lo, hi in self.bucket_ranges():      # NCCL picks algorithm / protocol by size: touch every bucket size once
                    hdist.allreduce_flat_grads(scratch[lo:hi])
                del scratch
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        grads = self._body_grads if self.n_streams <= 1 else (lambda: self._body_grads_multistream(self.n_streams))
        # Python's cyclic garbage collector must not run inside the capture: a finaliser of an unrelated CUDA object (an old
        # plan's graph, streams, events) that calls into the driver invalidates a global-mode stream capture
        import gc
        gc_was_on = gc.isenabled()
        gc.collect()
        gc.disable()
        try:
            if self.world > 1 and not self.ar_in_graph:
                with torch.cuda.graph(self.graph):
                    grads()
                self.graph_update = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_update):
                    self._body_update()
            else:
                with torch.cuda.graph(self.graph):
                    grads()
                    self._body_update()
        finally:
            if gc_was_on:
                gc.enable()
        # restore the state consumed by the warm-up step
        self.store.flat.copy_(saved[0])
        self.square_avg.copy_(saved[1])
        self.store.fbuf_flat.copy_(saved[2])
        self.store.ibuf_flat.copy_(saved[3])
        torch.cuda.synchronize(self.device)

    def close(self):
        """Drop the captured step graph(s).  With the all-reduce captured inside the graph this MUST happen before
        torch.distributed.destroy_process_group(): NCCL keeps a reference per captured collective and its communicator
        teardown waits until those graphs are gone."""
        if self.graph is not None or self.graph_update is not None:
            torch.cuda.synchronize(self.device)
            self.graph = None
            self.graph_update = None
            import gc
            gc.collect()

    def step(self, images, heatmaps):
        """Public end-to-end step: images [N,3,R,R], heatmaps [N,K,R/4,R/4] (host pinned or device).
        Returns the fp32 device scalar loss of this step (sum over stacks of per-stack MSE)."""
        self.x.copy_(images, non_blocking=True)
        self.t.copy_(heatmaps, non_blocking=True)
        return self.step_resident()

    # ---- double-buffered input path: the upload of batch i+1 overlaps the compute of batch i ----
    def prefetch(self, images, heatmaps):
        """Start the host->device copy of the NEXT batch (pinned host tensors, or device tensors) on a dedicated copy
        stream into staging buffers; returns immediately.  The batch is consumed by the next `step_prefetched()`.
        This is the `img.cuda(async=True)` of the reference loop (stack-hg.py:143-146) moved one iteration ahead, as a
        DataLoader with pin_memory does: PCIe traffic (25 MB per batch of 24) hides under the previous step."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self.xs, self.ts = torch.empty_like(self.x), torch.empty_like(self.t)
            self._staged = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream(self.device))
        cs = self._copy_stream
        cs.wait_event(self._staging_free)            # the previous staged batch has been moved into the static buffers
        with torch.cuda.stream(cs):
            # one copy per image: a single 19 MB DMA in flight next to the running step delayed the step's own launch
            # traffic on the same PCIe link by 0.3-1.2 ms (measured); sub-MB chunks leave gaps and cost nothing
            if images.dim() == 4 and images.shape[0] == self.xs.shape[0] and not images.is_cuda:
                for k in range(images.shape[0]):
                    self.xs[k].copy_(images[k], non_blocking=True)
                    self.ts[k].copy_(heatmaps[k], non_blocking=True)
            else:
                self.xs.copy_(images, non_blocking=True)
                self.ts.copy_(heatmaps, non_blocking=True)
            self._staged.record(cs)
        self._has_staged = True

    def step_prefetched(self, loss_to_host=False):
        """One train step on the batch uploaded by the last `prefetch()`; returns the fp32 device scalar loss.
        loss_to_host=True additionally queues the 4-byte device-to-host copy of this step's loss into a pinned slot;
        `pop_loss()` hands the losses out in order.  Reading step i's loss after step i+1 has been launched keeps the GPU
        busy across the host's read (a `loss.item()` right after the step leaves it idle for a launch latency every step)."""
        if not getattr(self, "_has_staged", False):
            raise HGKError("step_prefetched() without a preceding prefetch()")
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._staged)
        self.x.copy_(self.xs, non_blocking=True)     # device-to-device, ~10 us for 25 MB
        self.t.copy_(self.ts, non_blocking=True)
        self._staging_free.record(main)
        self._has_staged = False
        out = self.step_resident()
        if loss_to_host:
            if getattr(self, "_loss_pin", None) is None:
                self._loss_pin = torch.zeros(4, dtype=torch.float32).pin_memory()
                self._loss_ev = [torch.cuda.Event() for _ in range(4)]
                self._loss_q, self._loss_n = [], 0
            if len(self._loss_q) >= 4:
                raise HGKError("four losses are waiting in the read-back queue: call pop_loss()")
            slot = self._loss_n % 4
            self._loss_pin[slot:slot + 1].copy_(self.loss, non_blocking=True)
            self._loss_ev[slot].record(main)
            self._loss_q.append(slot)
            self._loss_n += 1
        return out

    def pop_loss(self):
        """The oldest loss queued by step_prefetched(loss_to_host=True) as a Python float (waits for that copy only)."""
        if not getattr(self, "_loss_q", None):
            raise HGKError("pop_loss(): no loss queued")
        slot = self._loss_q.pop(0)
        self._loss_ev[slot].synchronize()
        return float(self._loss_pin[slot])

    # ---- optimizer protocol of the reference loop (stack-hg.py:51-52,106; utils/checkpoint.py) ----
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        self.param_groups[0]["lr"] = value

    @property
    def alpha(self):
        return self.param_groups[0]["alpha"]

    @property
    def eps(self):
        return self.param_groups[0]["eps"]

    def state_dict(self):
        """checkpoint['optimizer'] of the reference: per-parameter square_avg + hyper-parameters (torch.optim layout)."""
        from .optim import pack_state
        return pack_state(self.store, self.square_avg, self.param_groups[0])

    def load_state_dict(self, sd):
        from .optim import unpack_state
        unpack_state(self.store, self.square_avg, sd, self.param_groups[0])

    def heatmaps(self):
        """Per-stack NCHW heat-maps of the last step (views of the plan's static output buffers)."""
        return [op.result for op in self.plan.outputs]
