"""The training step of train_model.py:939-960 as one re-playable unit.

``TrainStep`` owns static device buffers for one batch and runs ``optimizer.update(model, batch, itr)``
(train_model.py:950): forward, cleargrads, BPTT, gradient all-reduce (data parallel), Adam.  With ``graph=True`` the
whole device side -- ~1600 kernel launches for T=10 -- is captured once in a CUDA graph and replayed, which removes the
per-launch host cost (the strictly sequential ConvLSTM GEMMs are only 5-20 us each).  What changes from step to step
lives in memory the graph reads: the batch (static device buffers), the scheduled-sampling select (pinned host buffer
-> device copy node) and Adam's step counter (device int).
"""
import numpy as np
import torch

from ._lib import lib


class TrainStep(object):
    def __init__(self, model, optimizer, batch_size, seq_len, graph=True):
        self.model, self.opt = model, optimizer
        e = model.engine
        self.e = e
        dev = e.dev
        self.B, self.T = int(batch_size), int(seq_len)
        self.images = torch.zeros(self.T, self.B, 3, e.H, e.W, dtype=torch.float32, device=dev)
        self.actions = torch.zeros(self.T, self.B, 5, dtype=torch.float32, device=dev)
        self.states = torch.zeros(self.T, self.B, 5, dtype=torch.float32, device=dev)
        e._workspace(self.B, self.T)
        self.use_graph = bool(graph)
        self.graph = None
        self.launches_per_step = None
        self.feedself = (not model.train) or model.scheduled_sampling_k == -1

    # ---- data
    def load_batch(self, images, actions, states, non_blocking=True):
        """Copy one batch (host pinned or device tensors, time-major like concat_examples' output) into the static buffers."""
        self.images.copy_(images, non_blocking=non_blocking)
        self.actions.copy_(actions, non_blocking=non_blocking)
        self.states.copy_(states, non_blocking=non_blocking)

    # ---- one step
    def _fwd_bwd(self):
        e = self.e
        e.forward_device(self.images, self.actions, self.states, self.feedself)
        e.cleargrads()
        e.backward()

    def _update(self):
        """Fused 1/N scale + Adam on the flat buffers, then refresh of the bf16 weight operands (no collective in here)."""
        e, o = self.e, self.opt
        lib().call("pivp_adam_step", e.flat_p.data_ptr(), e.flat_g.data_ptr(), o.m.data_ptr(), o.v.data_ptr(), e.nparam,
                   o.step.data_ptr(), o.alpha, o.beta1, o.beta2, o.eps, 1.0 / self.model.world_size,
                   torch.cuda.current_stream(e.dev).cuda_stream)
        e.params_changed()

    def _overlap(self):
        """Data parallel, tensor-core mode, opt-in (PIVP_DP_OVERLAP=1): the gradient all-reduce runs unit by unit on a side stream under the
        deferred weight-gradient GEMMs (parallel.GradSync) and the whole step -- collective included -- is ONE CUDA graph.  Default: the
        plain form -- graph (forward + BPTT), one all-reduce of the flat buffer, graph (Adam).  Measured on 2 x B200 (b32 per GPU): 8.864 ms
        overlapped against 8.845 ms plain (8.750 ms on one GPU): the 36.8 MB all-reduce is ~0.1 ms over NVLink and the NCCL kernels take SMs
        from the GEMMs they run under, so nothing is gained at this size; and a process whose CUDA graphs hold NCCL kernels did not leave
        ``destroy_process_group`` within 5 minutes.  Correctness is covered either way (tests/test_gpu_dp.py ran with the overlap on)."""
        import os
        return self.model.world_size > 1 and self.e.compute == "bf16" and os.environ.get("PIVP_DP_OVERLAP", "0") == "1"

    def _eager_step(self):
        if self._overlap():
            from .parallel import GradSync
            if self.e.grad_sync is None:
                self.e.grad_sync = GradSync(self.e)
            self.e.grad_sync.begin()
            self._fwd_bwd()                        # backward() hands each finished unit to the side stream and joins it at the end
        else:
            self._fwd_bwd()
            if self.model.world_size > 1:
                from .parallel import allreduce_sum_
                allreduce_sum_(self.e.flat_g)      # the step's only collective: NCCL sum of 36.8 MB over NVLink
        self._update()

    def __call__(self, iter_num):
        m, e = self.model, self.e
        feedself, take, m.num_ground_truth = m.schedule(self.B * m.world_size, self.T, iter_num)
        assert feedself == self.feedself
        m.take_gt = take
        if not feedself:
            e.stage_schedule(take)
        if not self.use_graph:
            n0 = lib().query("pivp_launch_count")
            self._eager_step()
            self.launches_per_step = lib().query("pivp_launch_count") - n0
        else:
            if self.graph is None:
                self._capture()
            self.graph[0].replay()
            if self.graph[1] is not None:          # plain data-parallel form: the all-reduce sits between the two graphs, on the same stream
                from .parallel import allreduce_sum_
                allreduce_sum_(e.flat_g)
                self.graph[1].replay()
        m.gen_images = e.ws["gen"]
        m._bind_loss()
        return m.loss

    def _capture(self):
        """Capture the device side of the step.  Single GPU: one graph (forward + BPTT + Adam).  Data parallel: two graphs with
        the NCCL all-reduce launched between them, so no collective is recorded inside a graph."""
        dev = self.e.dev
        # one eager step on a side stream first: lazy attribute / module initialisation must not happen under capture.
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        saved = (self.e.flat_p.clone(), self.opt.m.clone(), self.opt.v.clone(), self.opt.step.clone())
        with torch.cuda.stream(s):
            self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.e.flat_p.copy_(saved[0]); self.opt.m.copy_(saved[1]); self.opt.v.copy_(saved[2]); self.opt.step.copy_(saved[3])
        self.e.params_changed()
        torch.cuda.synchronize(dev)
        n0 = lib().query("pivp_launch_count")
        ga = torch.cuda.CUDAGraph()
        gb = None
        if self._overlap():
            with torch.cuda.graph(ga):             # NCCL collectives on the forked side stream are captured with the kernels
                self._eager_step()
        elif self.model.world_size > 1:
            with torch.cuda.graph(ga):
                self._fwd_bwd()
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb):
                self._update()
        else:
            with torch.cuda.graph(ga):
                self._fwd_bwd()
                self._update()
        self.launches_per_step = lib().query("pivp_launch_count") - n0
        self.graph = (ga, gb)


class BatchPrefetcher(object):
    """Double-buffered host -> device staging in front of a TrainStep (the reference copies the batch synchronously at
    train_model.py:950, `xp.array`).  While step k computes, the batch of step k+1 travels from pinned host memory into a device
    staging buffer on a side stream; `load_next()` waits for that copy, moves it into the step's static input buffers with a
    stream-ordered device-to-device copy and starts the copy of the following batch.

        pf = BatchPrefetcher(step, batches)        # batches: iterable of (images, actions, states) pinned host tensors
        while pf.load_next():
            loss = step(iter_num)"""

    def __init__(self, step, batches):
        self.step = step
        dev = step.images.device
        self.stream = torch.cuda.Stream(device=dev)
        self.bufs = [tuple(torch.empty_like(x) for x in (step.images, step.actions, step.states)) for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.drained = [None, None]                # event behind the device-to-device copy that last read the slot
        self.ready = [False, False]
        self.it = iter(batches)
        self.k = 0
        self._issue(0)

    def _issue(self, slot):
        batch = next(self.it, None)
        self.ready[slot] = batch is not None
        if batch is None:
            return
        if self.drained[slot] is not None:         # wait for the copy that last read this slot -- not for the steps queued behind it
            self.stream.wait_event(self.drained[slot])
        with torch.cuda.stream(self.stream):
            for dst, src in zip(self.bufs[slot], batch):
                dst.copy_(src, non_blocking=True)
            self.events[slot].record(self.stream)

    def load_next(self, prefetch=True):
        """Put the next batch into the step's input buffers; False when the iterable is exhausted.  ``prefetch=False`` leaves the
        following batch un-requested until ``prefetch()`` is called: a caller whose batch source and scheduled-sampling both draw from
        the global NumPy stream (train_loop.py, like train_model.py:939-950) asks for batch k+1 only AFTER step k drew its selects, so
        the stream is consumed in the reference's order; the copy still overlaps step k on the device."""
        slot = self.k & 1
        if not self.ready[slot]:
            return False
        cur = torch.cuda.current_stream(self.stream.device)
        cur.wait_event(self.events[slot])
        self.step.load_batch(*self.bufs[slot])
        if self.drained[slot] is None:
            self.drained[slot] = torch.cuda.Event()
        self.drained[slot].record(cur)
        self.k += 1
        self.ready[self.k & 1] = False
        if prefetch:
            self._issue(self.k & 1)
        return True

    def prefetch(self):
        """Request the batch after the current one (see ``load_next(prefetch=False)``)."""
        if not self.ready[self.k & 1]:
            self._issue(self.k & 1)

