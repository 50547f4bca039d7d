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
    def _device_step(self):
        e = self.e
        e.forward_device(self.images, self.actions, self.states, self.feedself)
        e.cleargrads()
        e.backward()
        self.opt.apply()

    def __call__(self, iter_num):
        m, e = self.model, self.e
        feedself, take, m.num_ground_truth = m.schedule(self.B * m.world_size, self.T, iter_num)
        assert feedself == self.feedself
        m.take_gt = take
        if not feedself:
            e.stage_schedule(take)
        if not self.use_graph:
            n0 = lib().query("pivp_launch_count")
            self._device_step()
            self.launches_per_step = lib().query("pivp_launch_count") - n0
        else:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        m.gen_images = e.ws["gen"]
        m._bind_loss()
        return m.loss

    def _capture(self):
        # one eager step on a side stream first: lazy attribute / module initialisation must not happen under capture.
        # It is a real optimizer step; callers that need an untouched model capture on a scratch copy or reload parameters.
        s = torch.cuda.Stream(device=self.e.dev)
        s.wait_stream(torch.cuda.current_stream(self.e.dev))
        saved = (self.e.flat_p.clone(), self.opt.m.clone(), self.opt.v.clone(), self.opt.step.clone())
        with torch.cuda.stream(s):
            self._device_step()
        torch.cuda.current_stream(self.e.dev).wait_stream(s)
        torch.cuda.synchronize(self.e.dev)
        self.e.flat_p.copy_(saved[0]); self.opt.m.copy_(saved[1]); self.opt.v.copy_(saved[2]); self.opt.step.copy_(saved[3])
        self.e.params_changed()
        torch.cuda.synchronize(self.e.dev)
        g = torch.cuda.CUDAGraph()
        n0 = lib().query("pivp_launch_count")
        with torch.cuda.graph(g):
            self._device_step()
        self.launches_per_step = lib().query("pivp_launch_count") - n0
        self.graph = g
