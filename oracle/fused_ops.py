"""Array-in / array-out oracles for the fused transform + mask-softmax + composite ops.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function rebuilds, with the
``npgrad`` operators, exactly the sub-graph of the reference between the outputs of the
1x1 deconvolutions / kernel Linear and ``gen_images[t]`` (the boundary SURVEY 8d uses for
the algorithmic byte count), and returns the forward value plus a ``bwd(g_out)`` closure
giving the gradients w.r.t. every input.  Reference lines (``src/models/train_model.py``):
CDNA :315-317,326-349 ; DNA :388-415 ; STP :454-471 ; masks + composite :719-728.
"""
import numpy as np
from . import npgrad as G
from .model import RELU_SHIFT, DNA_KERN_SIZE, _broadcast_div, _broadcast_scale


def _masks_and_composite(prev, layers, a_pre, M):
    """ref:719-728: relu -> reshape(-1, M+1) on NCHW memory -> softmax -> composite (zip-truncated)."""
    B, _, H, W = prev.shape
    m = G.softmax(G.reshape(G.relu(a_pre), (-1, M + 1)))
    m = G.reshape(m, (B, M + 1, H, W))
    ml = G.split_axis(m, M + 1, 1)
    out = _broadcast_scale(prev, ml[0])
    for layer, mask in zip(layers, ml[1:]):
        out = out + _broadcast_scale(layer, mask)
    return out, m


def _finish(out, m, inputs, extra=None):
    def bwd(g_out):
        for v in inputs.values():
            v.grad = None
        G.backward(out, seed=np.asarray(g_out, out.data.dtype))
        return {k: (np.zeros_like(v.data) if v.grad is None else v.grad) for k, v in inputs.items()}
    res = dict(out=out.data, masks=m.data, bwd=bwd)
    if extra:
        res.update(extra)
    return res


def cdna_fused(prev, enc7_pre, mask_pre, kern_raw, num_masks):
    """prev (B,3,H,W), enc7_pre (B,3,H,W), mask_pre (B,M+1,H,W), kern_raw (B,25*M) -> out (B,3,H,W)."""
    M = num_masks
    iv = dict(prev=G.Var(prev), enc7_pre=G.Var(enc7_pre), mask_pre=G.Var(mask_pre), kern_raw=G.Var(kern_raw))
    B, C, H, W = prev.shape
    layers = [G.sigmoid(G.relu(iv["enc7_pre"]))]
    k = G.reshape(iv["kern_raw"], (B, M, 1, DNA_KERN_SIZE, DNA_KERN_SIZE))
    k = G.relu(k - RELU_SHIFT) + RELU_SHIFT
    k = _broadcast_div(k, G.sum_(k, (2, 3, 4), keepdims=True))
    kn = k
    k = G.transpose(G.reshape(k, (B, M, DNA_KERN_SIZE, DNA_KERN_SIZE)), (1, 0, 2, 3))
    t = G.depthwise_convolution_2d(G.transpose(iv["prev"], (1, 0, 2, 3)), k, 1, DNA_KERN_SIZE // 2)
    t = G.transpose(G.reshape(t, (C, B, M, H, W)), (2, 1, 0, 3, 4))
    layers += [G.squeeze(s, 0) for s in G.split_axis(t, M, 0)]
    out, m = _masks_and_composite(iv["prev"], layers, iv["mask_pre"], M)
    return _finish(out, m, iv, dict(kern_norm=kn.data.reshape(B, M, 25), transformed=t.data))


def dna_fused(prev, enc7_pre, mask_pre):
    """prev (B,3,H,W), enc7_pre (B,25,H,W), mask_pre (B,2,H,W) -> out (B,3,H,W); taps detached (B.2)."""
    iv = dict(prev=G.Var(prev), enc7_pre=G.Var(enc7_pre), mask_pre=G.Var(mask_pre))
    B, C, H, W = prev.shape
    enc7 = G.relu(iv["enc7_pre"])
    padded = np.pad(prev, ((0, 0), (0, 0), (2, 2), (2, 2)), mode="constant")
    taps = []
    for xk in range(DNA_KERN_SIZE):
        for yk in range(DNA_KERN_SIZE):
            win = np.pad(padded[:, :, xk:H, yk:W], ((0, 0), (0, 0), (0, xk), (0, yk)), mode="constant")
            taps.append(win[:, None])
    taps = G.Var(np.concatenate(taps, axis=1))
    k = G.relu(enc7 - RELU_SHIFT) + RELU_SHIFT
    k = _broadcast_div(k, G.sum_(k, 1, keepdims=True))
    t = G.sum_(_broadcast_scale(taps, G.expand_dims(k, 2)), 1)
    out, m = _masks_and_composite(iv["prev"], [t], iv["mask_pre"], 1)
    return _finish(out, m, iv, dict(transformed=t.data))


def stp_fused(prev, enc7_pre, mask_pre, theta, num_masks, oob="zeros"):
    """prev, enc7_pre (B,3,H,W), mask_pre (B,M+1,H,W), theta (B,6) (identity already added) -> out.

    The M-1 transformers share one Linear (B.4) so they all sample with the same ``theta``;
    the gradient returned for ``theta`` is the SUM over the M-1 uses, as autograd produces.
    """
    M = num_masks
    iv = dict(prev=G.Var(prev), enc7_pre=G.Var(enc7_pre), mask_pre=G.Var(mask_pre), theta=G.Var(theta))
    B, C, H, W = prev.shape
    layers = [G.sigmoid(iv["enc7_pre"])]
    for _ in range(M - 1):
        grid = G.spatial_transformer_grid(G.reshape(iv["theta"], (B, 2, 3)), (H, W))
        layers.append(G.spatial_transformer_sampler(iv["prev"], grid, oob))
    out, m = _masks_and_composite(iv["prev"], layers, iv["mask_pre"], M)
    return _finish(out, m, iv, dict(warped=layers[1].data if M > 1 else None))


def resize_images(x, out_hw):
    """chainer.functions.resize_images (Chainer 2.0.1 ResizeImages.forward) as predict_model.py:120 calls it: bilinear, sample positions
    ``linspace(0, W-1, out_W)`` in float64, corner indices clipped to ``[0, W-2]``, weights cast to the input dtype."""
    B, C, H, W = x.shape
    out_H, out_W = out_hw
    u_1d = np.linspace(0, W - 1, num=out_W)
    v_1d = np.linspace(0, H - 1, num=out_H)
    grid = np.meshgrid(u_1d, v_1d)
    u, v = grid[0].ravel(), grid[1].ravel()
    u0 = np.floor(u).astype(np.int32).clip(0, W - 2)
    v0 = np.floor(v).astype(np.int32).clip(0, H - 2)
    u1, v1 = u0 + 1, v0 + 1
    w1 = ((u1 - u) * (v1 - v)).astype(x.dtype)
    w2 = ((u - u0) * (v1 - v)).astype(x.dtype)
    w3 = ((u1 - u) * (v - v0)).astype(x.dtype)
    w4 = ((u - u0) * (v - v0)).astype(x.dtype)
    y = (w1[None, None, :] * x[:, :, v0, u0] + w2[None, None, :] * x[:, :, v0, u1] +
         w3[None, None, :] * x[:, :, v1, u0] + w4[None, None, :] * x[:, :, v1, u1])
    return y.reshape(B, C, out_H, out_W)
