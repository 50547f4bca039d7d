"""Parameter layouts: Chainer parameter paths / shapes <-> the private kernel layouts.

The conversions are pure permutations (exact), so a reference ``.npz`` checkpoint written by
``serializers.save_npz`` (train_model.py:1035) loads bit-exactly and gradients can be compared with the
oracle in Chainer layout.  All parameters live in ONE flat fp32 buffer (so the gradient all-reduce and the
Adam update are single flat operations); this module fixes the order and offsets.
"""
import numpy as np

LSTM_SIZES = (32, 32, 64, 64, 128, 64, 32)        # train_model.py:509-515
LSTM_IN = (32, 32, 32, 64, 64, 128, 96)           # channels of the layer input x (App. C)
DNA_KERN_SIZE = 5


def gate_perm(C):
    """Chainer row r = gate*C + ch  ->  internal row (ch//32)*128 + gate*32 + ch%32 (gate order j,i,f,o; A.10)."""
    r = np.arange(4 * C)
    gate, ch = r // C, r % C
    return (ch // 32) * 128 + gate * 32 + ch % 32


class ParamSpec(object):
    def __init__(self, name, chainer_shape, kind, extra=None):
        self.name, self.chainer_shape, self.kind, self.extra = name, tuple(chainer_shape), kind, extra
        self.size = int(np.prod(chainer_shape))
        self.offset = None

    # ---- Chainer -> internal
    def to_internal(self, a):
        a = np.asarray(a, np.float32).reshape(self.chainer_shape)
        k = self.kind
        if k == "conv":                       # (O,I,kh,kw) -> (O,kh,kw,I)
            return np.ascontiguousarray(a.transpose(0, 2, 3, 1))
        if k == "deconv":                     # (in,out,kh,kw) -> (in,kh,kw,out)
            return np.ascontiguousarray(a.transpose(0, 2, 3, 1))
        if k == "head":                       # 1x1 deconv (in,out,1,1) -> (out,in)
            return np.ascontiguousarray(a[:, :, 0, 0].T)
        if k == "lstm_w":                     # rows permuted to the gate-interleaved order, then OHWI
            out = np.empty_like(a)
            out[gate_perm(self.extra)] = a
            return np.ascontiguousarray(out.transpose(0, 2, 3, 1))
        if k == "lstm_b":
            out = np.empty_like(a)
            out[gate_perm(self.extra)] = a
            return out
        if k == "ln":                         # (C*H*W,) -> (H*W*C,)
            C, H, W = self.extra
            return np.ascontiguousarray(a.reshape(C, H, W).transpose(1, 2, 0)).reshape(-1)
        if k == "linear_chw":                 # (out, C*H*W) -> (out, H*W*C)
            C, H, W = self.extra
            return np.ascontiguousarray(a.reshape(-1, C, H, W).transpose(0, 2, 3, 1)).reshape(a.shape[0], -1)
        return np.ascontiguousarray(a)

    # ---- internal -> Chainer
    def to_chainer(self, a):
        a = np.asarray(a, np.float32)
        k = self.kind
        s = self.chainer_shape
        if k in ("conv", "deconv"):
            return np.ascontiguousarray(a.reshape(s[0], s[2], s[3], s[1]).transpose(0, 3, 1, 2))
        if k == "head":
            return np.ascontiguousarray(a.reshape(s[1], s[0]).T).reshape(s)
        if k == "lstm_w":
            a = a.reshape(s[0], s[2], s[3], s[1]).transpose(0, 3, 1, 2)
            return np.ascontiguousarray(a[gate_perm(self.extra)])
        if k == "lstm_b":
            return np.ascontiguousarray(a.reshape(-1)[gate_perm(self.extra)])
        if k == "ln":
            C, H, W = self.extra
            return np.ascontiguousarray(a.reshape(H, W, C).transpose(2, 0, 1)).reshape(-1)
        if k == "linear_chw":
            C, H, W = self.extra
            return np.ascontiguousarray(a.reshape(-1, H, W, C).transpose(0, 3, 1, 2)).reshape(s)
        return np.ascontiguousarray(a.reshape(s))


def param_specs(model_type, num_masks, use_state, H, W):
    """Ordered list of ParamSpec with flat offsets.  Order groups the two 1x1 heads so they form one matrix."""
    h2, w2, h4, w4, h8, w8 = H // 2, W // 2, H // 4, W // 4, H // 8, W // 8
    sa = 10 if use_state else 0
    sp = []

    def add(name, shape, kind="plain", extra=None):
        sp.append(ParamSpec(name, shape, kind, extra))

    add("enc0/W", (32, 3, 5, 5), "conv"); add("enc0/b", (32,))
    add("enc1/W", (32, 32, 3, 3), "conv"); add("enc1/b", (32,))
    add("enc2/W", (64, 64, 3, 3), "conv"); add("enc2/b", (64,))
    add("enc3/W", (64, 64 + sa, 1, 1), "conv"); add("enc3/b", (64,))
    add("enc4/W", (128, 128, 3, 3), "deconv"); add("enc4/b", (128,))
    add("enc5/W", (96, 96, 3, 3), "deconv"); add("enc5/b", (96,))
    add("enc6/W", (64, 64, 3, 3), "deconv"); add("enc6/b", (64,))
    for i, (cin, c) in enumerate(zip(LSTM_IN, LSTM_SIZES), 1):
        add("lstm%d/conv/W" % i, (4 * c, cin + c, 5, 5), "lstm_w", c)
        add("lstm%d/conv/b" % i, (4 * c,), "lstm_b", c)
    ln = [("norm_enc0", (32, h2, w2)), ("norm_enc6", (64, H, W)), ("hidden1", (32, h2, w2)), ("hidden2", (32, h2, w2)),
          ("hidden3", (64, h4, w4)), ("hidden4", (64, h4, w4)), ("hidden5", (128, h8, w8)), ("hidden6", (64, h4, w4)),
          ("hidden7", (32, h2, w2))]
    for name, chw in ln:
        n = chw[0] * chw[1] * chw[2]
        add(name + "/norm/gamma", (n,), "ln", chw)
        add(name + "/norm/beta", (n,), "ln", chw)
    ne = {"CDNA": 3, "DNA": DNA_KERN_SIZE ** 2, "STP": 3}[model_type]
    # heads: enc7 then masks, weights adjacent and biases adjacent -> one (ne+M+1, 64) matrix
    add("model/enc7/W", (64, ne, 1, 1), "head"); add("masks/W", (64, num_masks + 1, 1, 1), "head")
    add("model/enc7/b", (ne,)); add("masks/b", (num_masks + 1,))
    add("current_state/W", (5, 10)); add("current_state/b", (5,))
    if model_type == "CDNA":
        add("model/cdna_kerns/W", (DNA_KERN_SIZE ** 2 * num_masks, 128 * h8 * w8), "linear_chw", (128, h8, w8))
        add("model/cdna_kerns/b", (DNA_KERN_SIZE ** 2 * num_masks,))
    elif model_type == "STP":
        add("model/stp_input/W", (100, 128 * h8 * w8), "linear_chw", (128, h8, w8))
        add("model/stp_input/b", (100,))
        add("model/identity_params/W", (6, 100)); add("model/identity_params/b", (6,))
    # every tensor starts 16-byte aligned, except the second member of a head pair, which must follow
    # its partner without a gap (the pair is addressed as one matrix / one bias vector)
    glued = ("masks/W", "masks/b")
    off = 0
    for i, s in enumerate(sp):
        if s.name not in glued:
            off = (off + 3) // 4 * 4
        s.offset = off
        off += s.size
    return sp, (off + 3) // 4 * 4


def lecun_normal_init(specs, seed=4321):
    """Chainer defaults (A.1): W ~ N(0, 1/fan_in), fan_in = prod(W.shape[1:]); b = 0; gamma = 1, beta = 0.
    Drawn with RandomState(seed) in sorted param-path order -> identical to oracle.model.init_params."""
    rs = np.random.RandomState(seed)
    out = {}
    for s in sorted(specs, key=lambda q: q.name):
        if s.name.endswith("/W"):
            fan_in = int(np.prod(s.chainer_shape[1:]))
            out[s.name] = (rs.standard_normal(s.chainer_shape) * np.sqrt(1.0 / fan_in)).astype(np.float32)
        elif s.name.endswith("gamma"):
            out[s.name] = np.ones(s.chainer_shape, np.float32)
        else:
            out[s.name] = np.zeros(s.chainer_shape, np.float32)
    return out
