"""Forward-only inference rollout, as ``predict_model.py:99-128`` uses the model.

The reference resizes the raw frames with ``F.resize_images`` and ``/ 255.0`` (:118-123), then calls
``model([imgs, acts, stas], 0)`` under ``chainer.using_config('train', False)`` (:126-128): ``feedself`` is on
(train_model.py:649-650), the first ``context_frames`` steps see the ground truth, every later step its own prediction, and
``model.gen_images`` is read back.  Batch size is 1 there.

``Rollout`` owns static device buffers and ONE captured CUDA graph of the forward pass (no backward, no optimizer), so a
prediction is: resize kernel -> graph replay -> read ``gen_images``.  The tensor-core (bf16) path tiles the 8x8 maps of ConvLSTM 5
in pairs of images, so an odd batch (the reference's batch of 1) is padded with a copy of its last sequence; every operation of
the model is per-sample, hence the padding changes no output and (a duplicated sample contributes its own error again) not the
batch-mean loss of a batch of one either.
"""
import numpy as np
import torch

from ._lib import lib


class Rollout(object):
    def __init__(self, model, batch_size, seq_len, graph=True):
        self.model, self.e = model, model.engine
        e = self.e
        self.B, self.T = int(batch_size), int(seq_len)
        self.Bp = self.B + (self.B & 1) if e.compute == "bf16" else self.B
        dev = e.dev
        self.images = torch.zeros(self.T, self.Bp, 3, e.H, e.W, dtype=torch.float32, device=dev)
        self.actions = torch.zeros(self.T, self.Bp, 5, dtype=torch.float32, device=dev)
        self.states = torch.zeros(self.T, self.Bp, 5, dtype=torch.float32, device=dev)
        self.use_graph, self.graph = bool(graph), None

    # ---- inputs
    def _put(self, dst, src):
        src = torch.as_tensor(src)
        dst[:, :self.B].copy_(src.to(dst.device, non_blocking=True))
        if self.Bp > self.B:
            dst[:, self.B:].copy_(dst[:, self.B - 1:self.B])

    def load(self, images, actions, states):
        """images (T,B,3,H,W) float32 already at the model's size and in [0,1]; actions / states (T,B,5)."""
        self._put(self.images, images); self._put(self.actions, actions); self._put(self.states, states)

    def load_raw(self, raw_images, actions, states, divide_by=255.0):
        """predict_model.py:118-123: raw frames (T,B,3,H0,W0), float32 or uint8 -> bilinear resize to the model's size and / 255."""
        e = self.e
        raw = torch.as_tensor(raw_images)
        if raw.dtype not in (torch.uint8, torch.float32):
            raw = raw.float()
        raw = raw.to(e.dev).contiguous()
        T, B, C, H0, W0 = raw.shape
        assert (T, B, C) == (self.T, self.B, 3)
        out = torch.empty(T, B, 3, e.H, e.W, dtype=torch.float32, device=e.dev)
        lib().call("pivp_resize_images", raw.data_ptr(), 1 if raw.dtype == torch.uint8 else 0, out.data_ptr(), T * B * 3, H0, W0, e.H, e.W,
                   float(divide_by), 1, torch.cuda.current_stream(e.dev).cuda_stream)
        self.load(out, actions, states)

    # ---- the rollout
    def _forward(self):
        self.e.forward_device(self.images, self.actions, self.states, True)      # feedself (train_model.py:649-650, 664-666)

    def __call__(self, images=None, actions=None, states=None):
        """Returns the list of T-1 predicted frames, each (B,3,H,W); ``model.loss`` / ``psnr_all`` / ``gen_images`` are set as in the reference."""
        m, e = self.model, self.e
        if images is not None:
            self.load(images, actions, states)
        e._workspace(self.Bp, self.T)
        if not self.use_graph:
            self._forward()
        else:
            if self.graph is None:
                dev = e.dev
                s = torch.cuda.Stream(device=dev)
                s.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(s):
                    self._forward()                                           # lazy initialisation outside the capture
                torch.cuda.current_stream(dev).wait_stream(s)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._forward()
            self.graph.replay()
        m.num_ground_truth, m.take_gt = None, None
        m.gen_images = [g[:self.B] for g in e.ws["gen"][:self.T - 1]]
        m._bind_loss()
        return m.gen_images


def predict(model, raw_images, actions, states, divide_by=255.0):
    """The compute part of predict_model.py:main (:99-128) for one or more sequences given like ``concat_examples`` returns them
    (time-major): resize + /255, test-mode rollout, predicted frames back as a NumPy array (T-1, B, 3, H, W)."""
    T, B = int(np.shape(raw_images)[0]), int(np.shape(raw_images)[1])
    was = model.train
    model.train = False                                   # chainer.using_config('train', False)
    try:
        r = Rollout(model, B, T, graph=False)
        r.load_raw(raw_images, actions, states, divide_by)
        gen = r()
        out = torch.stack(list(gen)).cpu().numpy()
    finally:
        model.train = was
    return out
