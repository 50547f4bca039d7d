"""Tiny reverse-mode autograd over NumPy with Chainer-2.0.1 operator semantics.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every operator states the
Chainer function it restates and the reference call sites that use it
(``src/models/train_model.py`` line numbers).  Algorithms follow Chainer's CPU
path: im2col + tensordot for convolutions, one NumPy call per elementary op.

The graph is dynamic: every op returns a ``Var`` holding ``data`` and a closure
that maps the output gradient to the parents' gradients.  ``backward(loss)``
walks the graph in reverse topological order, like ``Variable.backward``.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view


class Var(object):
    """Array + graph node (stands in for ``chainer.Variable``)."""
    __slots__ = ("data", "grad", "parents", "bwd", "name", "saved_grad")

    def __init__(self, data, parents=(), bwd=None, name=None):
        self.data = data
        self.grad = None
        self.parents = parents
        self.bwd = bwd
        self.name = name

    @property
    def shape(self):
        return self.data.shape

    # arithmetic sugar (Chainer Variables overload the same operators)
    def __add__(self, o):
        return add(self, o)
    __radd__ = __add__

    def __sub__(self, o):
        return add(self, neg(o) if isinstance(o, Var) else -o)

    def __mul__(self, o):
        return mul(self, o)
    __rmul__ = __mul__

    def __truediv__(self, o):
        return div(self, o)


def as_var(x):
    return x if isinstance(x, Var) else Var(np.asarray(x))


def detach(x):
    """``Variable(x.data)`` / ``.data`` -- cuts the graph (train_model.py:404,670)."""
    return Var(as_var(x).data)


KEEP_INTERMEDIATE = False


def backward(loss, seed=None):
    """Reverse sweep (``loss.backward()`` inside ``optimizer.update``, train_model.py:950)."""
    order, seen = [], set()
    stack = [(loss, False)]
    while stack:
        v, done = stack.pop()
        if done:
            order.append(v)
            continue
        if id(v) in seen:
            continue
        seen.add(id(v))
        stack.append((v, True))
        for p in v.parents:
            if id(p) not in seen:
                stack.append((p, False))
    loss.grad = np.ones_like(loss.data) if seed is None else seed
    for v in reversed(order):
        if v.bwd is None or v.grad is None:
            continue
        gs = v.bwd(v.grad)
        for p, g in zip(v.parents, gs):
            if g is None:
                continue
            p.grad = g if p.grad is None else p.grad + g
        if v.parents:
            if KEEP_INTERMEDIATE:
                v.saved_grad = v.grad                    # debugging aid (scripts/dbg_bwd_stage.py): keep d(loss)/d(intermediate)
            v.grad = None if v is not loss else v.grad   # free intermediate grads


# ----------------------------------------------------------------------------
# elementwise / shape ops (NumPy semantics)
# ----------------------------------------------------------------------------

def _unbroadcast(g, shape):
    if g.shape == shape:
        return g
    nd = g.ndim - len(shape)
    if nd:
        g = g.sum(axis=tuple(range(nd)))
    ax = tuple(i for i, s in enumerate(shape) if s == 1 and g.shape[i] != 1)
    if ax:
        g = g.sum(axis=ax, keepdims=True)
    return g


def add(a, b):
    a = as_var(a)
    if not isinstance(b, Var):
        return Var(a.data + b, (a,), lambda g: (g,))
    return Var(a.data + b.data, (a, b),
               lambda g: (_unbroadcast(g, a.data.shape), _unbroadcast(g, b.data.shape)))


def neg(a):
    return Var(-a.data, (a,), lambda g: (-g,))


def mul(a, b):
    a = as_var(a)
    if not isinstance(b, Var):
        return Var(a.data * b, (a,), lambda g: (g * b,))
    return Var(a.data * b.data, (a, b),
               lambda g: (_unbroadcast(g * b.data, a.data.shape),
                          _unbroadcast(g * a.data, b.data.shape)))


def div(a, b):
    a = as_var(a)
    if not isinstance(b, Var):
        return Var(a.data / b, (a,), lambda g: (g / b,))
    y = a.data / b.data
    return Var(y, (a, b),
               lambda g: (_unbroadcast(g / b.data, a.data.shape),
                          _unbroadcast(-g * y / b.data, b.data.shape)))


def relu(x):
    """``F.relu`` (train_model.py:316,327,388,408,459,698,719)."""
    y = np.maximum(x.data, 0)
    return Var(y, (x,), lambda g: (g * (y > 0),))


def sigmoid(x):
    """``F.sigmoid`` (train_model.py:271-272,317,455): tanh(x/2)/2+1/2 like Chainer's CPU path."""
    half = x.data.dtype.type(0.5)
    y = np.tanh(x.data * half) * half + half
    return Var(y, (x,), lambda g: (g * y * (1 - y),))


def tanh(x):
    """``F.tanh`` (train_model.py:271-272)."""
    y = np.tanh(x.data)
    return Var(y, (x,), lambda g: (g * (1 - y * y),))


def log(x):
    """``F.log`` (train_model.py:134)."""
    return Var(np.log(x.data), (x,), lambda g: (g / x.data,))


def reshape(x, shape):
    """``F.reshape``: C-order view of the NCHW buffer (this is what makes B.1 happen)."""
    s0 = x.data.shape
    return Var(x.data.reshape(shape), (x,), lambda g: (g.reshape(s0),))


def transpose(x, axes):
    inv = np.argsort(axes)
    return Var(x.data.transpose(axes), (x,), lambda g: (g.transpose(inv),))


def concat(xs, axis=1):
    """``F.concat`` (train_model.py:262,565,575,676)."""
    xs = [as_var(x) for x in xs]
    sizes = [x.data.shape[axis] for x in xs]
    cuts = np.cumsum(sizes)[:-1]
    return Var(np.concatenate([x.data for x in xs], axis=axis), tuple(xs),
               lambda g: tuple(np.split(g, cuts, axis=axis)))


def split_axis(x, sections, axis):
    """``F.split_axis`` into equal sections (train_model.py:269,346,723)."""
    n = x.data.shape[axis] // sections
    outs = []
    for i in range(sections):
        sl = [slice(None)] * x.data.ndim
        sl[axis] = slice(i * n, (i + 1) * n)
        outs.append(getitem(x, tuple(sl)))
    return outs


def getitem(x, sl):
    def bwd(g):
        gx = np.zeros_like(x.data)
        gx[sl] = g
        return (gx,)
    return Var(x.data[sl], (x,), bwd)


def squeeze(x, axis):
    s0 = x.data.shape
    return Var(np.squeeze(x.data, axis=axis), (x,), lambda g: (g.reshape(s0),))


def expand_dims(x, axis):
    s0 = x.data.shape
    return Var(np.expand_dims(x.data, axis), (x,), lambda g: (g.reshape(s0),))


def broadcast_to(x, shape):
    """``F.broadcast_to``; backward sums over the broadcast axes (A.6)."""
    s0 = x.data.shape
    return Var(np.broadcast_to(x.data, shape), (x,), lambda g: (_unbroadcast(g, s0),))


def tile(x, reps):
    """``F.tile`` (train_model.py:564)."""
    s0 = x.data.shape

    def bwd(g):
        # fold every tiled axis back: reshape to (rep, size) pairs and sum the reps
        shp, ax = [], []
        for i, (r, s) in enumerate(zip(reps, s0)):
            shp += [r, s]
            ax.append(2 * i)
        return (g.reshape(shp).sum(axis=tuple(ax)),)
    return Var(np.tile(x.data, reps), (x,), bwd)


def sum_(x, axis=None, keepdims=False):
    """``F.sum`` (train_model.py:328,409,414)."""
    s0 = x.data.shape

    def bwd(g):
        if not keepdims and axis is not None:
            g = np.expand_dims(g, axis)
        return (np.broadcast_to(g, s0).astype(x.data.dtype, copy=False),)
    return Var(x.data.sum(axis=axis, keepdims=keepdims), (x,), bwd)


def pad_const(x, pad_width):
    """``F.pad(mode='constant', constant_values=0)`` (train_model.py:395,402)."""
    sl = tuple(slice(lo, lo + s) for (lo, _), s in zip(pad_width, x.data.shape))
    return Var(np.pad(x.data, pad_width, mode="constant"), (x,), lambda g: (g[sl],))


def softmax(x):
    """``F.softmax`` axis=1, max-subtracted (train_model.py:721; A.6)."""
    e = np.exp(x.data - x.data.max(axis=1, keepdims=True))
    y = e / e.sum(axis=1, keepdims=True)

    def bwd(g):
        gx = y * g
        gx -= y * gx.sum(axis=1, keepdims=True)
        return (gx,)
    return Var(y, (x,), bwd)


def mean_squared_error(a, b):
    """``F.mean_squared_error``: mean over ALL elements (train_model.py:134,741,751; A.6)."""
    a, b = as_var(a), as_var(b)
    d = a.data - b.data
    n = d.dtype.type(d.size)
    y = np.asarray((d * d).sum() / n, dtype=d.dtype)

    def bwd(g):
        gd = (2 * g / n) * d
        return (gd, -gd)
    return Var(y, (a, b), bwd)


# ----------------------------------------------------------------------------
# connection ops (im2col + tensordot, Chainer CPU algorithm class; A.2-A.5)
# ----------------------------------------------------------------------------

def im2col(x, kh, kw, s, p, out_h=None, out_w=None):
    """(B,C,H,W) -> (B,C,kh,kw,Ho,Wo) view-copy, zero padding (Chainer ``im2col_cpu``)."""
    B, C, H, W = x.shape
    if out_h is None:
        out_h = (H + 2 * p - kh) // s + 1
        out_w = (W + 2 * p - kw) // s + 1
    # pad enough on the bottom/right for the last window (matters when outsize is forced, A.3)
    need_h = (out_h - 1) * s + kh
    need_w = (out_w - 1) * s + kw
    xp = np.pad(x, ((0, 0), (0, 0), (p, max(p, need_h - H - p)), (p, max(p, need_w - W - p))),
                mode="constant")
    win = sliding_window_view(xp, (kh, kw), axis=(2, 3))      # (B,C,H',W',kh,kw)
    win = win[:, :, ::s, ::s][:, :, :out_h, :out_w]
    return win.transpose(0, 1, 4, 5, 2, 3)


def col2im(col, s, p, H, W):
    """Adjoint of ``im2col`` (Chainer ``col2im_cpu``): (B,C,kh,kw,Ho,Wo) -> (B,C,H,W)."""
    B, C, kh, kw, Ho, Wo = col.shape
    img = np.zeros((B, C, H + 2 * p + s - 1, W + 2 * p + s - 1), dtype=col.dtype)
    for i in range(kh):
        for j in range(kw):
            img[:, :, i:i + s * Ho:s, j:j + s * Wo:s] += col[:, :, i, j]
    return img[:, :, p:p + H, p:p + W]


def convolution_2d(x, W, b, stride, pad):
    """``L.Convolution2D`` / ``F.convolution_2d``: cross-correlation, cover_all=False (A.2).

    Call sites: train_model.py:224 (ConvLSTM gates), :500-503 (enc0..enc3).
    """
    kh, kw = W.data.shape[2:]
    col = im2col(x.data, kh, kw, stride, pad)
    y = np.tensordot(col, W.data, ((1, 2, 3), (1, 2, 3)))      # (B,Ho,Wo,O)
    y += b.data
    y = np.ascontiguousarray(np.rollaxis(y, 3, 1))
    H, Wd = x.data.shape[2:]

    def bwd(gy):
        gW = np.tensordot(gy, col, ((0, 2, 3), (0, 4, 5))).astype(W.data.dtype, copy=False)
        gcol = np.tensordot(W.data, gy, (0, 1))                # (C,kh,kw,B,Ho,Wo)
        gx = col2im(np.rollaxis(gcol, 3), stride, pad, H, Wd)
        return (gx, gW, gy.sum(axis=(0, 2, 3)))
    return Var(y, (x, W, b), bwd)


def deconvolution_2d(x, W, b, stride, pad, outsize):
    """``L.Deconvolution2D`` with explicit ``outsize`` (A.3).  W is (in,out,kh,kw).

    Call sites: train_model.py:505-507 (enc4..enc6), :288,364,429 (enc7), :527 (masks).
    """
    kh, kw = W.data.shape[2:]
    oh, ow = outsize
    B, _, ih, iw = x.data.shape
    assert (oh + 2 * pad - kh) // stride + 1 == ih and (ow + 2 * pad - kw) // stride + 1 == iw
    gcol = np.tensordot(W.data, x.data, (0, 1))                # (out,kh,kw,B,ih,iw)
    y = col2im(np.rollaxis(gcol, 3), stride, pad, oh, ow)
    y = y + b.data.reshape(1, -1, 1, 1)

    def bwd(gy):
        col = im2col(gy, kh, kw, stride, pad, ih, iw)          # (B,out,kh,kw,ih,iw)
        gW = np.tensordot(x.data, col, ((0, 2, 3), (0, 4, 5))).astype(W.data.dtype, copy=False)
        gx = np.tensordot(col, W.data, ((1, 2, 3), (1, 2, 3)))  # (B,ih,iw,in)
        gx = np.ascontiguousarray(np.rollaxis(gx, 3, 1))
        return (gx, gW, gy.sum(axis=(0, 2, 3)))
    return Var(y, (x, W, b), bwd)


def linear(x, W, b):
    """``L.Linear``: y = x W^T + b, W (out,in) (train_model.py:289,430,431,529)."""
    y = x.data.dot(W.data.T) + b.data

    def bwd(gy):
        return (gy.dot(W.data), gy.T.dot(x.data), gy.sum(axis=0))
    return Var(y, (x, W, b), bwd)


def layer_normalization(x, gamma, beta, eps=1e-6):
    """``L.LayerNormalization`` on a (B, n) view: biased variance, per-element gamma/beta (A.4).

    Call site: train_model.py:192,203-208 (nine instances).
    """
    mu = x.data.mean(axis=1, keepdims=True)
    xc = x.data - mu
    var = (xc * xc).mean(axis=1, keepdims=True)
    rstd = 1 / np.sqrt(var + x.data.dtype.type(eps))
    xh = xc * rstd
    y = xh * gamma.data + beta.data

    def bwd(gy):
        q = gy * gamma.data
        gx = (q - q.mean(axis=1, keepdims=True) - xh * (q * xh).mean(axis=1, keepdims=True)) * rstd
        return (gx, (gy * xh).sum(axis=0), gy.sum(axis=0))
    return Var(y, (x, gamma, beta), bwd)


def depthwise_convolution_2d(x, W, stride, pad):
    """``F.depthwise_convolution_2d(x (N,C,H,W), W (D,C,kh,kw))`` -> (N, C*D, H', W') (A.5).

    Output channel index is ``c*D + d``.  Call site: train_model.py:341 with N=3 colours,
    C=batch, D=num_masks.  CPU algorithm: im2col then a per-channel matmul.
    """
    D, C, kh, kw = W.data.shape
    N = x.data.shape[0]
    col = im2col(x.data, kh, kw, stride, pad)                  # (N,C,kh,kw,Ho,Wo)
    Ho, Wo = col.shape[4:]
    y = np.einsum("ncijhw,dcij->ncdhw", col, W.data, optimize=True)
    H, Wd = x.data.shape[2:]

    def bwd(gy):
        g = gy.reshape(N, C, D, Ho, Wo)
        gW = np.einsum("ncdhw,ncijhw->dcij", g, col, optimize=True).astype(W.data.dtype, copy=False)
        gcol = np.einsum("ncdhw,dcij->ncijhw", g, W.data, optimize=True)
        return (col2im(gcol, stride, pad, H, Wd), gW)
    return Var(y.reshape(N, C * D, Ho, Wo), (x, W), bwd)


# ----------------------------------------------------------------------------
# spatial transformer (A.7)
# ----------------------------------------------------------------------------

def spatial_transformer_grid(theta, out_hw):
    """``F.spatial_transformer_grid(theta (B,2,3), (H,W))`` -> (B,2,H,W) (train_model.py:469).

    grid[b,:,i,j] = theta[b] . [x_j, y_i, 1]^T with x,y = linspace(-1,1,.); channel 0 is x.
    """
    H, W = out_hw
    dt = theta.data.dtype
    xs = np.linspace(-1, 1, W, dtype=dt)
    ys = np.linspace(-1, 1, H, dtype=dt)
    base = np.stack([np.broadcast_to(xs[None, :], (H, W)),
                     np.broadcast_to(ys[:, None], (H, W)),
                     np.ones((H, W), dtype=dt)]).reshape(3, H * W)
    B = theta.data.shape[0]
    y = theta.data.dot(base).reshape(B, 2, H, W) if theta.data.ndim == 2 else \
        np.einsum("bij,jk->bik", theta.data, base).reshape(B, 2, H, W)

    def bwd(g):
        return (np.einsum("bik,jk->bij", g.reshape(B, 2, H * W), base).astype(dt, copy=False),)
    return Var(y, (theta,), bwd)


def spatial_transformer_sampler(x, grid, oob="zeros"):
    """``F.spatial_transformer_sampler``: bilinear sampling at u_pix=(u+1)(W-1)/2 (train_model.py:470).

    ``oob`` selects the out-of-range rule that A.7 leaves open for Chainer 2.0.1's CPU code:
    "zeros"  -- corners outside the image contribute 0 (cuDNN path, later Chainer, torch
                ``grid_sample(padding_mode='zeros', align_corners=True)``);
    "border" -- coordinates are clipped to the image first (early Chainer CPU code); the
                clipped coordinates receive no gradient.
    """
    X, G = x.data, grid.data
    B, C, H, W = X.shape
    dt = X.dtype
    u = (G[:, 0] + 1) * dt.type((W - 1) / 2.0)                 # (B,Ho,Wo) pixel x
    v = (G[:, 1] + 1) * dt.type((H - 1) / 2.0)
    live_u = np.ones_like(u, dtype=bool)
    live_v = np.ones_like(v, dtype=bool)
    if oob == "border":
        live_u = (u >= 0) & (u <= W - 1)
        live_v = (v >= 0) & (v <= H - 1)
        u = np.clip(u, 0, W - 1)
        v = np.clip(v, 0, H - 1)
    u0 = np.floor(u)
    v0 = np.floor(v)
    fu = (u - u0).astype(dt)
    fv = (v - v0).astype(dt)
    u0 = u0.astype(np.int64)
    v0 = v0.astype(np.int64)
    bidx = np.arange(B)[:, None, None]

    def corner(vi, ui):
        ok = (vi >= 0) & (vi < H) & (ui >= 0) & (ui < W)
        val = X[bidx, :, np.clip(vi, 0, H - 1), np.clip(ui, 0, W - 1)]   # (B,Ho,Wo,C)
        return val * ok[..., None], ok

    p00, k00 = corner(v0, u0)
    p01, k01 = corner(v0, u0 + 1)
    p10, k10 = corner(v0 + 1, u0)
    p11, k11 = corner(v0 + 1, u0 + 1)
    w00 = ((1 - fv) * (1 - fu))[..., None]
    w01 = ((1 - fv) * fu)[..., None]
    w10 = (fv * (1 - fu))[..., None]
    w11 = (fv * fu)[..., None]
    y = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11).transpose(0, 3, 1, 2)

    def bwd(gy):
        g = gy.transpose(0, 2, 3, 1)                            # (B,Ho,Wo,C)
        gx = np.zeros_like(X)
        gxt = gx.transpose(0, 2, 3, 1)                          # view (B,H,W,C)
        for (vi, ui, w, ok) in ((v0, u0, w00, k00), (v0, u0 + 1, w01, k01),
                                (v0 + 1, u0, w10, k10), (v0 + 1, u0 + 1, w11, k11)):
            contrib = g * w * ok[..., None]
            np.add.at(gxt, (np.broadcast_to(bidx, vi.shape), np.clip(vi, 0, H - 1),
                            np.clip(ui, 0, W - 1)), contrib)
        # d/du, d/dv of the bilinear form (corners outside contribute value 0)
        du = ((p01 - p00) * (1 - fv)[..., None] + (p11 - p10) * fv[..., None])
        dv = ((p10 - p00) * (1 - fu)[..., None] + (p11 - p01) * fu[..., None])
        gu = (g * du).sum(axis=3) * dt.type((W - 1) / 2.0) * live_u
        gv = (g * dv).sum(axis=3) * dt.type((H - 1) / 2.0) * live_v
        return (gx, np.stack([gu, gv], axis=1).astype(dt, copy=False))
    return Var(np.ascontiguousarray(y), (x, grid), bwd)
