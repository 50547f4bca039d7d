"""NumPy restatement of the reference training step.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows ``/root/reference/src/models/train_model.py`` (cited as ``ref:<line>``):
helpers ref:51-180, LayerNormalizationConv2D ref:186-208, BasicConvLSTMCell ref:216-276,
StatelessCDNA ref:278-351, StatelessDNA ref:354-417, StatelessSTP ref:419-475,
Model ref:478-764, Adam step ref:860-861,950-960.  The reference's behaviours that
differ from the TF original (flat-11 softmax grouping, truncated/detached DNA taps,
zip truncation, shared STP Linear ...) are reproduced on purpose.

The model is expressed functionally: ``params`` is a flat dict keyed by Chainer
parameter paths (the keys ``serializers.save_npz`` writes, ref:1035), the
recurrent state lives in a local dict, and ``forward`` returns everything the
parity tests compare.  The only generalisation over the reference is the image
size: deconvolution ``outsize`` is (H/4, H/2, H) instead of the literals 16/32/64
(ref:505-507) so that small test cases and the 128x128 STP configuration run.
"""
import math
import numpy as np
from . import npgrad as G

RELU_SHIFT = 1e-12        # ref:42
DNA_KERN_SIZE = 5         # ref:45
LSTM_SIZES = (32, 32, 64, 64, 128, 64, 32)   # ref:509-515


class Config(object):
    """Constructor arguments of ``Model`` (ref:484) plus image geometry."""

    def __init__(self, model_type="CDNA", num_masks=10, use_state=True, schedsamp_k=-1.0,
                 context_frames=2, height=64, width=64, train=True, stp_oob="zeros",
                 dtype=np.float32):
        assert model_type in ("CDNA", "DNA", "STP")
        assert height % 8 == 0 and width % 8 == 0
        self.model_type = model_type
        self.num_masks = num_masks
        self.use_state = use_state
        self.schedsamp_k = schedsamp_k
        self.context_frames = context_frames
        self.height, self.width = height, width
        self.train = train
        self.stp_oob = stp_oob
        self.dtype = np.dtype(dtype)


# ----------------------------------------------------------------------------
# parameters (A.1, A.9) and synthetic data
# ----------------------------------------------------------------------------

def param_shapes(cfg):
    """Chainer param path -> shape, with the lazily inferred in-sizes filled in (A.1, App. C)."""
    H, W = cfg.height, cfg.width
    s2, s4, s8 = (H // 2) * (W // 2), (H // 4) * (W // 4), (H // 8) * (W // 8)
    sa = 10 if cfg.use_state else 0
    shp = {
        "enc0/W": (32, 3, 5, 5), "enc1/W": (32, 32, 3, 3), "enc2/W": (64, 64, 3, 3),
        "enc3/W": (64, 64 + sa, 1, 1),
        "enc4/W": (128, 128, 3, 3), "enc5/W": (96, 96, 3, 3), "enc6/W": (64, 64, 3, 3),  # deconv: (in,out,k,k)
        "masks/W": (64, cfg.num_masks + 1, 1, 1),
        "current_state/W": (5, 10),
    }
    lstm_in = (32, 32, 32, 64, 64, 128, 96)
    for i, (cin, c) in enumerate(zip(lstm_in, LSTM_SIZES), 1):
        shp["lstm%d/conv/W" % i] = (4 * c, cin + c, 5, 5)
    ln = {"norm_enc0": 32 * s2, "norm_enc6": 64 * H * W, "hidden1": 32 * s2, "hidden2": 32 * s2,
          "hidden3": 64 * s4, "hidden4": 64 * s4, "hidden5": 128 * s8, "hidden6": 64 * s4,
          "hidden7": 32 * s2}
    for k, n in ln.items():
        shp[k + "/norm/gamma"] = (n,)
        shp[k + "/norm/beta"] = (n,)
    if cfg.model_type == "CDNA":
        shp["model/enc7/W"] = (64, 3, 1, 1)
        shp["model/cdna_kerns/W"] = (DNA_KERN_SIZE ** 2 * cfg.num_masks, 128 * s8)
    elif cfg.model_type == "DNA":
        shp["model/enc7/W"] = (64, DNA_KERN_SIZE ** 2, 1, 1)
    else:
        shp["model/enc7/W"] = (64, 3, 1, 1)
        shp["model/stp_input/W"] = (100, 128 * s8)
        shp["model/identity_params/W"] = (6, 100)
    for k in list(shp):
        if k.endswith("/W"):
            w = shp[k]
            is_deconv = k.split("/")[-2] in ("enc4", "enc5", "enc6", "masks", "enc7")
            shp[k[:-1] + "b"] = (w[1] if is_deconv else w[0],)
    return shp


def init_params(cfg, seed=4321):
    """Chainer defaults: W ~ LeCunNormal = N(0, 1/fan_in), fan_in = prod(W.shape[1:]); b=0; LN gamma=1, beta=0.

    Drawn with ``RandomState(seed)`` in sorted param-path order.
    """
    rs = np.random.RandomState(seed)
    out = {}
    for k, s in sorted(param_shapes(cfg).items()):
        if k.endswith("/W"):
            fan_in = int(np.prod(s[1:]))
            out[k] = (rs.standard_normal(s) * math.sqrt(1.0 / fan_in)).astype(cfg.dtype)
        elif k.endswith("gamma"):
            out[k] = np.ones(s, cfg.dtype)
        else:
            out[k] = np.zeros(s, cfg.dtype)
    return out


def synthetic_sequences(batch, seq_len, cfg, seed=1234, blobs=True):
    """Push-style synthetic sequences in the on-disk layout of make_dataset.py:104-136.

    Returns a list of ``[image (T,H,W,3) in [0,1], action (T,5), state (T,5)]`` float32.
    """
    rs = np.random.RandomState(seed)
    H, W = cfg.height, cfg.width
    seqs = []
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for _ in range(batch):
        if blobs:
            img = np.zeros((seq_len, H, W, 3), np.float32) + rs.rand(1, 1, 1, 3).astype(np.float32) * 0.3
            for _b in range(3):
                p = rs.rand(2) * (H, W)
                vel = rs.uniform(-2, 2, 2)
                col = rs.rand(3).astype(np.float32)
                sig = rs.uniform(2, 6)
                for t in range(seq_len):
                    c = p + vel * t
                    img[t] += np.exp(-((yy - c[0]) ** 2 + (xx - c[1]) ** 2) / (2 * sig * sig))[..., None] * col
            img += rs.rand(seq_len, H, W, 3).astype(np.float32) * 0.05
            img = np.clip(img, 0, 1).astype(np.float32)
        else:
            img = rs.rand(seq_len, H, W, 3).astype(np.float32)
        act = rs.uniform(-1, 1, (seq_len, 5)).astype(np.float32)
        sta = rs.uniform(-1, 1, (seq_len, 5)).astype(np.float32)
        seqs.append([img, act, sta])
    return seqs


def concat_examples(batch):
    """ref:51-71: list of [img (T,H,W,3), act (T,5), sta (T,5)] -> time-major channel-first arrays."""
    img = np.array([b[0] for b in batch])
    act = np.array([b[1] for b in batch])
    sta = np.array([b[2] for b in batch])
    img = np.ascontiguousarray(img.transpose(1, 0, 4, 2, 3))    # split on axis 1 + rollaxis(.,3,1)
    return img, np.ascontiguousarray(act.transpose(1, 0, 2)), np.ascontiguousarray(sta.transpose(1, 0, 2))


# ----------------------------------------------------------------------------
# scheduled sampling (ref:73-122, 649-657) -- integer/index work, bit-exact gate
# ----------------------------------------------------------------------------

def num_ground_truth(batch_size, k, iter_num):
    """ref:653-655: int32(round(float32(B) * (k / (k + exp(iter/k))))), round-half-even."""
    return np.int32(np.round(np.float32(batch_size) * (k / (k + np.exp(iter_num / k)))))


def scheduled_sample_order(batch_size, n_gt, rng=np.random):
    """The index work of ref:93-96: one legacy ``shuffle(arange(B))``; first n_gt entries pick ground truth.

    Returns a bool array ``take_gt[b]``.  ref:98-121 (stack/argsort/stitch) reduce to this select.
    """
    idx = np.arange(int(batch_size))
    rng.shuffle(idx)
    take = np.zeros(int(batch_size), dtype=bool)
    take[idx[:int(n_gt)]] = True
    return take


def scheduled_sample(ground_truth_x, generated_x, batch_size, n_gt, rng=np.random):
    """ref:73-122 restated literally (index bookkeeping included) -- used to prove the select form."""
    idx = np.arange(int(batch_size))
    rng.shuffle(idx)
    gt_idx = idx[:int(n_gt)]
    gen_idx = idx[int(n_gt):]
    gt_flat = ground_truth_x.reshape(int(batch_size), -1)
    gen_flat = generated_x.reshape(int(batch_size), -1)
    gt_ex = np.take(gt_flat, gt_idx, axis=0)
    gen_ex = np.take(gen_flat, gen_idx, axis=0)
    tags = np.hstack((np.vstack((gt_idx, np.zeros_like(gt_idx))),
                      np.vstack((gen_idx, np.ones_like(gen_idx)))))
    tags = tags[:, np.argsort(np.hstack((gt_idx, gen_idx)))]
    rows = []
    for i in range(tags.shape[1]):
        if tags[1][i] == 0:
            rows.append(gt_ex[np.where(gt_idx == i)])
        else:
            rows.append(gen_ex[np.where(gen_idx == i)])
    return np.array(rows, dtype=np.float32).reshape(ground_truth_x.shape)


# ----------------------------------------------------------------------------
# sub-links
# ----------------------------------------------------------------------------

def layer_norm_conv2d(P, name, x):
    """ref:203-208: flatten (B,C,H,W)->(B,CHW), LayerNormalization(eps=1e-6), reshape back."""
    s = x.shape
    y = G.layer_normalization(G.reshape(x, (s[0], -1)), P[name + "/norm/gamma"], P[name + "/norm/beta"])
    return G.reshape(y, s)


def conv_lstm(P, name, x, state, out_size, forget_bias=1.0):
    """ref:234-276.  Gate order j,i,f,o (A.10); recurrent state is the raw h (B.6)."""
    if state.get(name) is None:
        z = np.zeros((x.shape[0], out_size, x.shape[2], x.shape[3]), x.data.dtype)   # ref:254-257
        state[name] = (G.Var(z), G.Var(z.copy()))
    c, h = state[name]
    gates = G.convolution_2d(G.concat((x, h), axis=1), P[name + "/conv/W"], P[name + "/conv/b"], 1, 2)
    j, i, f, o = G.split_axis(gates, 4, 1)
    c = c * G.sigmoid(f + forget_bias) + G.sigmoid(i) * G.tanh(j)
    h = G.tanh(c) * G.sigmoid(o)
    state[name] = (c, h)
    return h


def _broadcast_div(x, y):
    """ref:152-165 ``broadcasted_division``: materialised broadcast then divide."""
    return x / G.broadcast_to(y, x.shape)


def _broadcast_scale(x, y):
    """ref:167-180 ``broadcast_scale``."""
    return x * G.broadcast_to(y, x.shape)


def cdna_transform(P, cfg, enc6, hidden5, prev_image, trace):
    """StatelessCDNA.__call__ ref:311-351."""
    B, C, H, W = prev_image.shape
    M = cfg.num_masks
    e_pre = G.deconvolution_2d(enc6, P["model/enc7/W"], P["model/enc7/b"], 1, 0, (H, W))
    enc7 = G.relu(e_pre)
    out = [G.sigmoid(enc7)]                                              # ref:315-317
    r = G.linear(G.reshape(hidden5, (B, -1)), P["model/cdna_kerns/W"], P["model/cdna_kerns/b"])
    k = G.reshape(r, (B, M, 1, DNA_KERN_SIZE, DNA_KERN_SIZE))
    k = G.relu(k - RELU_SHIFT) + RELU_SHIFT                             # ref:327
    k = _broadcast_div(k, G.sum_(k, (2, 3, 4), keepdims=True))          # ref:328-329
    k = G.transpose(G.reshape(k, (B, M, DNA_KERN_SIZE, DNA_KERN_SIZE)), (1, 0, 2, 3))
    p = G.transpose(prev_image, (1, 0, 2, 3))                           # colours become the batch axis
    t = G.depthwise_convolution_2d(p, k, 1, DNA_KERN_SIZE // 2)         # (3, B*M, H, W)  ref:341
    t = G.transpose(G.reshape(t, (C, B, M, H, W)), (2, 1, 0, 3, 4))     # ref:344-345
    out += [G.squeeze(s, 0) for s in G.split_axis(t, M, 0)]
    trace.update(enc7_pre=e_pre, kern_raw=r, kern_norm=k)
    return out, enc7


def dna_transform(P, cfg, enc6, prev_image, trace):
    """StatelessDNA.__call__ ref:384-417 (taps truncated at H / W and detached, B.2)."""
    B, C, H, W = prev_image.shape
    if cfg.num_masks != 1:
        raise ValueError("Only one mask is supported for DNA model.")   # ref:389-390
    e_pre = G.deconvolution_2d(enc6, P["model/enc7/W"], P["model/enc7/b"], 1, 0, (H, W))
    enc7 = G.relu(e_pre)
    padded = np.pad(prev_image.data, ((0, 0), (0, 0), (2, 2), (2, 2)), mode="constant")
    taps = []
    for xk in range(DNA_KERN_SIZE):
        for yk in range(DNA_KERN_SIZE):
            win = padded[:, :, xk:H, yk:W]                               # ref:400 (upper bound H, not H+4)
            win = np.pad(win, ((0, 0), (0, 0), (0, xk), (0, yk)), mode="constant")   # ref:402
            taps.append(win[:, None])                                    # .data => detached, ref:404
    taps = G.Var(np.concatenate(taps, axis=1))                           # (B,25,3,H,W)
    k = G.relu(enc7 - RELU_SHIFT) + RELU_SHIFT                           # ref:408
    k = _broadcast_div(k, G.sum_(k, 1, keepdims=True))                   # ref:409-410
    k = G.expand_dims(k, 2)
    t = G.sum_(_broadcast_scale(taps, k), 1)                             # ref:413-414
    trace.update(enc7_pre=e_pre)
    return [t], enc7


def stp_transform(P, cfg, enc6, hidden5, prev_image, trace):
    """StatelessSTP.__call__ ref:449-475 (no ReLU on enc7; ONE Linear(6) shared by all transformers, B.4)."""
    B, C, H, W = prev_image.shape
    enc7 = G.deconvolution_2d(enc6, P["model/enc7/W"], P["model/enc7/b"], 1, 0, (H, W))
    out = [G.sigmoid(enc7)]
    s = G.relu(G.linear(G.reshape(hidden5, (B, -1)), P["model/stp_input/W"], P["model/stp_input/b"]))
    ident = np.tile(np.array([[1, 0, 0, 0, 1, 0]], enc7.data.dtype), (B, 1))
    thetas = []
    for _ in range(cfg.num_masks - 1):
        th = G.linear(s, P["model/identity_params/W"], P["model/identity_params/b"]) + ident
        th = G.reshape(th, (B, 2, 3))
        grid = G.spatial_transformer_grid(th, (H, W))
        out.append(G.spatial_transformer_sampler(prev_image, grid, cfg.stp_oob))
        thetas.append(th)
    trace.update(enc7_pre=enc7, theta=thetas[0] if thetas else None)
    return out, enc7


# ----------------------------------------------------------------------------
# Model.__call__ (ref:620-764)
# ----------------------------------------------------------------------------

def forward(params, batch, iter_num, cfg, rng=np.random, take_gt_log=None):
    """One forward pass over a time-major batch ``(images (T,B,3,H,W), actions (T,B,5), states (T,B,5))``.

    Returns a dict: ``loss`` (Var), ``psnr_all``, ``gen_images`` / ``gen_states`` (lists of Var),
    ``P`` (param Vars, whose ``.grad`` is filled by ``npgrad.backward(loss)``) and ``trace``
    (per-step intermediates the kernel parity tests compare).
    """
    images, actions, states = batch
    dt = cfg.dtype
    P = {k: G.Var(np.asarray(v, dt), name=k) for k, v in params.items()}
    T, B = images.shape[0], images.shape[1]
    H, W = cfg.height, cfg.width
    ctx = cfg.context_frames
    state = {}
    gen_images, gen_states, trace = [], [], []
    current_state = G.Var(np.asarray(states[0], dt))

    feedself = (not cfg.train) or cfg.schedsamp_k == -1                  # ref:649
    n_gt = None if feedself else num_ground_truth(B, cfg.schedsamp_k, iter_num)

    for t in range(T - 1):                                               # ref:659
        tr = {}
        image = np.asarray(images[t], dt)
        done_warm_start = len(gen_images) > ctx - 1                      # ref:663
        if feedself and done_warm_start:
            prev_image = gen_images[-1]                                  # keeps the graph, ref:666
        elif done_warm_start:
            take = scheduled_sample_order(B, n_gt, rng)                  # ref:670 (detached)
            if take_gt_log is not None:
                take_gt_log.append(take.copy())
            prev_image = G.Var(np.where(take[:, None, None, None], image, gen_images[-1].data))
        else:
            prev_image = G.Var(image)                                    # ref:673
        action = G.Var(np.asarray(actions[t], dt))
        state_action = G.concat((action, current_state), axis=1)        # ref:676

        # ops table ref:594-602, ReLU after every group ref:698
        encs = []
        x = layer_norm_conv2d(P, "norm_enc0", G.convolution_2d(prev_image, P["enc0/W"], P["enc0/b"], 2, 2))
        encs.append(G.relu(x))
        h1 = layer_norm_conv2d(P, "hidden1", conv_lstm(P, "lstm1", encs[0], state, 32))
        h2 = layer_norm_conv2d(P, "hidden2", conv_lstm(P, "lstm2", h1, state, 32))
        encs.append(G.relu(G.convolution_2d(h2, P["enc1/W"], P["enc1/b"], 2, 1)))
        h3 = layer_norm_conv2d(P, "hidden3", conv_lstm(P, "lstm3", encs[1], state, 64))
        h4 = layer_norm_conv2d(P, "hidden4", conv_lstm(P, "lstm4", h3, state, 64))
        encs.append(G.relu(G.convolution_2d(h4, P["enc2/W"], P["enc2/b"], 2, 1)))
        x = encs[2]
        if cfg.use_state:                                                # ref:559-566
            smear = G.reshape(state_action, (B, 10, 1, 1))
            smear = G.tile(smear, (1, 1, x.shape[2], x.shape[3]))
            x = G.concat((x, smear), axis=1)
        encs.append(G.relu(G.convolution_2d(x, P["enc3/W"], P["enc3/b"], 1, 0)))
        h5 = layer_norm_conv2d(P, "hidden5", conv_lstm(P, "lstm5", encs[3], state, 128))
        encs.append(G.relu(G.deconvolution_2d(h5, P["enc4/W"], P["enc4/b"], 2, 1, (H // 4, W // 4))))
        h6 = layer_norm_conv2d(P, "hidden6", conv_lstm(P, "lstm6", encs[4], state, 64))
        x = G.concat((h6, encs[1]), axis=1)                              # ref:600
        encs.append(G.relu(G.deconvolution_2d(x, P["enc5/W"], P["enc5/b"], 2, 1, (H // 2, W // 2))))
        h7 = layer_norm_conv2d(P, "hidden7", conv_lstm(P, "lstm7", encs[5], state, 32))
        x = G.concat((h7, encs[0]), axis=1)                              # ref:601
        x = layer_norm_conv2d(P, "norm_enc6", G.deconvolution_2d(x, P["enc6/W"], P["enc6/b"], 2, 1, (H, W)))
        encs.append(G.relu(x))
        enc6 = encs[6]

        if cfg.model_type == "CDNA":
            transformed, enc7 = cdna_transform(P, cfg, enc6, h5, prev_image, tr)
        elif cfg.model_type == "DNA":
            transformed, enc7 = dna_transform(P, cfg, enc6, prev_image, tr)
        else:
            transformed, enc7 = stp_transform(P, cfg, enc6, h5, prev_image, tr)

        # masks: softmax over groups of (num_masks+1) FLAT NCHW elements (ref:718-723, B.1)
        a_pre = G.deconvolution_2d(enc6, P["masks/W"], P["masks/b"], 1, 0, (H, W))
        m = G.softmax(G.reshape(G.relu(a_pre), (-1, cfg.num_masks + 1)))
        m = G.reshape(m, (B, cfg.num_masks + 1, H, W))
        mask_list = G.split_axis(m, cfg.num_masks + 1, 1)
        output = _broadcast_scale(prev_image, mask_list[0])             # ref:725
        for layer, mask in zip(transformed, mask_list[1:]):             # zip truncation, B.3
            output = output + _broadcast_scale(layer, mask)
        gen_images.append(output)

        current_state = G.linear(state_action, P["current_state/W"], P["current_state/b"])   # ref:730
        gen_states.append(current_state)
        tr.update(prev_image=prev_image, enc6=enc6, hidden5=h5, mask_pre=a_pre, masks=m,
                  transformed=transformed, output=output, encs=encs,
                  hiddens=[h1, h2, h3, h4, h5, h6, h7], state_action=state_action)
        trace.append(tr)

    # loss ref:737-758
    loss, psnr_all, recon = 0.0, 0.0, []
    for x, gx in zip(images[ctx:], gen_images[ctx - 1:]):
        c = G.mean_squared_error(G.Var(np.asarray(x, dt)), gx)
        recon.append(float(c.data))
        psnr_all = psnr_all + 10.0 * math.log(1.0 / float(c.data)) / math.log(10.0)   # ref:134
        loss = c + loss
    for s, gs in zip(states[ctx:], gen_states[ctx - 1:]):
        loss = G.mean_squared_error(G.Var(np.asarray(s, dt)), gs) * 1e-4 + loss       # ref:751
    loss = loss / dt.type(T - ctx)                                       # ref:758
    return dict(loss=loss, psnr_all=psnr_all, recon_costs=recon, gen_images=gen_images,
                gen_states=gen_states, P=P, trace=trace, n_gt=n_gt)


def loss_and_grads(params, batch, iter_num, cfg, rng=np.random):
    """forward + ``cleargrads`` + ``backward`` (the first three lines of ``Optimizer.update``, A.8)."""
    out = forward(params, batch, iter_num, cfg, rng)
    G.backward(out["loss"])
    grads = {}
    for k, v in out["P"].items():
        grads[k] = np.zeros_like(v.data) if v.grad is None else v.grad   # None grads are zero-filled (A.8)
    out["grads"] = grads
    return out


class Adam(object):
    """Chainer 2.0.1 ``AdamRule`` (A.8): eps is added to sqrt(v) un-corrected; t starts at 1."""

    def __init__(self, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def lr(self):
        return self.alpha * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)

    def update(self, params, grads):
        self.t += 1
        lr = self.lr()
        for k, p in params.items():
            g = grads[k]
            m = self.m.setdefault(k, np.zeros_like(p))
            v = self.v.setdefault(k, np.zeros_like(p))
            m += (1 - self.beta1) * (g - m)
            v += (1 - self.beta2) * (g * g - v)
            p -= (lr * m / (np.sqrt(v) + self.eps)).astype(p.dtype)


def train_step(params, adam, batch, iter_num, cfg, rng=np.random):
    """ref:950-960: ``optimizer.update(model, [imgs, acts, stas], itr)`` then ``reset_state``."""
    out = loss_and_grads(params, batch, iter_num, cfg, rng)
    adam.update(params, out["grads"])
    return out
