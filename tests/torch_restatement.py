"""Independent torch-CPU (autograd) restatement of the reference model -- used ONLY to pin the oracle.

It deliberately uses different primitives than ``oracle/`` (``F.conv2d``,
``conv_transpose2d(output_padding=1)``, ``F.layer_norm``, grouped conv for the CDNA
transform, ``grid_sample``) so that an error in the oracle's hand-written
im2col/col2im/backward code cannot cancel out.  Reference lines: see ``oracle/model.py``.
"""
import math
import torch
import torch.nn.functional as F

EPS = 1e-12


def _ln(P, name, x):
    s = x.shape
    n = s[1] * s[2] * s[3]
    return F.layer_norm(x.reshape(s[0], n), (n,), P[name + "/norm/gamma"], P[name + "/norm/beta"], 1e-6).reshape(s)


def _lstm(P, name, x, st, C):
    if name not in st:
        z = x.new_zeros(x.shape[0], C, x.shape[2], x.shape[3])
        st[name] = (z, z)
    c, h = st[name]
    g = F.conv2d(torch.cat([x, h], 1), P[name + "/conv/W"], P[name + "/conv/b"], padding=2)
    j, i, f, o = torch.chunk(g, 4, 1)
    c = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
    h = torch.tanh(c) * torch.sigmoid(o)
    st[name] = (c, h)
    return h


def _deconv(P, name, x, stride):
    if stride == 1:
        return F.conv_transpose2d(x, P[name + "/W"], P[name + "/b"])
    return F.conv_transpose2d(x, P[name + "/W"], P[name + "/b"], stride=2, padding=1, output_padding=1)


def forward(params, batch, cfg, take_gt=None, dtype=torch.float64):
    """``take_gt``: list of bool arrays (one per scheduled-sampling step) or None for feedself."""
    images, actions, states = [torch.as_tensor(a, dtype=dtype) for a in batch]
    P = {k: torch.tensor(v, dtype=dtype, requires_grad=True) for k, v in params.items()}
    T, B = images.shape[:2]
    H, W = cfg.height, cfg.width
    M = cfg.num_masks
    ctx = cfg.context_frames
    st = {}
    gen_images, gen_states, masks_out = [], [], []
    cur = states[0]
    n_sched = 0
    for t in range(T - 1):
        warm = len(gen_images) > ctx - 1
        if warm and take_gt is None:
            prev = gen_images[-1]
        elif warm:
            sel = torch.as_tensor(take_gt[n_sched])[:, None, None, None]
            n_sched += 1
            prev = torch.where(sel, images[t], gen_images[-1].detach())
        else:
            prev = images[t]
        sa = torch.cat([actions[t], cur], 1)
        e0 = F.relu(_ln(P, "norm_enc0", F.conv2d(prev, P["enc0/W"], P["enc0/b"], stride=2, padding=2)))
        h1 = _ln(P, "hidden1", _lstm(P, "lstm1", e0, st, 32))
        h2 = _ln(P, "hidden2", _lstm(P, "lstm2", h1, st, 32))
        e1 = F.relu(F.conv2d(h2, P["enc1/W"], P["enc1/b"], stride=2, padding=1))
        h3 = _ln(P, "hidden3", _lstm(P, "lstm3", e1, st, 64))
        h4 = _ln(P, "hidden4", _lstm(P, "lstm4", h3, st, 64))
        e2 = F.relu(F.conv2d(h4, P["enc2/W"], P["enc2/b"], stride=2, padding=1))
        x = e2
        if cfg.use_state:
            x = torch.cat([x, sa[:, :, None, None].expand(B, 10, x.shape[2], x.shape[3])], 1)
        e3 = F.relu(F.conv2d(x, P["enc3/W"], P["enc3/b"]))
        h5 = _ln(P, "hidden5", _lstm(P, "lstm5", e3, st, 128))
        e4 = F.relu(_deconv(P, "enc4", h5, 2))
        h6 = _ln(P, "hidden6", _lstm(P, "lstm6", e4, st, 64))
        e5 = F.relu(_deconv(P, "enc5", torch.cat([h6, e1], 1), 2))
        h7 = _ln(P, "hidden7", _lstm(P, "lstm7", e5, st, 32))
        e6 = F.relu(_ln(P, "norm_enc6", _deconv(P, "enc6", torch.cat([h7, e0], 1), 2)))

        if cfg.model_type == "CDNA":
            layers = [torch.sigmoid(F.relu(_deconv(P, "model/enc7", e6, 1)))]
            k = F.linear(h5.reshape(B, -1), P["model/cdna_kerns/W"], P["model/cdna_kerns/b"]).reshape(B, M, 25)
            k = F.relu(k - EPS) + EPS
            k = k / k.sum(2, keepdim=True)
            # per-sample kernels: batch on the group axis
            inp = prev.permute(1, 0, 2, 3)                                   # (3,B,H,W)
            tr = F.conv2d(inp, k.reshape(B * M, 1, 5, 5), padding=2, groups=B)   # (3,B*M,H,W)
            tr = tr.reshape(3, B, M, H, W)
            layers += [tr[:, :, m].permute(1, 0, 2, 3) for m in range(M)]
        elif cfg.model_type == "DNA":
            e7 = F.relu(_deconv(P, "model/enc7", e6, 1))
            k = F.relu(e7 - EPS) + EPS
            k = k / k.sum(1, keepdim=True)
            pd = prev.detach()
            out = 0
            for xk in range(5):
                for yk in range(5):
                    # tap[i,j] = prev[i+xk-2, j+yk-2] if 2 <= i+xk < H and 2 <= j+yk < W else 0   (B.2)
                    tap = torch.zeros_like(pd)
                    i0, i1 = max(0, 2 - xk), H - xk
                    j0, j1 = max(0, 2 - yk), W - yk
                    tap[:, :, i0:i1, j0:j1] = pd[:, :, i0 + xk - 2:i1 + xk - 2, j0 + yk - 2:j1 + yk - 2]
                    out = out + k[:, xk * 5 + yk][:, None] * tap
            layers = [out]
        else:
            layers = [torch.sigmoid(_deconv(P, "model/enc7", e6, 1))]
            s = F.relu(F.linear(h5.reshape(B, -1), P["model/stp_input/W"], P["model/stp_input/b"]))
            ident = torch.tensor([1.0, 0, 0, 0, 1, 0], dtype=dtype)
            for _ in range(M - 1):
                th = (F.linear(s, P["model/identity_params/W"], P["model/identity_params/b"]) + ident).reshape(B, 2, 3)
                grid = F.affine_grid(th, (B, 3, H, W), align_corners=True)
                layers.append(F.grid_sample(prev, grid, mode="bilinear", align_corners=True,
                                            padding_mode="zeros" if cfg.stp_oob == "zeros" else "border"))

        a = F.relu(_deconv(P, "masks", e6, 1))
        m = torch.softmax(a.reshape(-1, M + 1), 1).reshape(B, M + 1, H, W)
        out = prev * m[:, 0:1]
        for layer, k_ in zip(layers, range(1, M + 1)):
            out = out + layer * m[:, k_:k_ + 1]
        gen_images.append(out)
        masks_out.append(m)
        cur = F.linear(sa, P["current_state/W"], P["current_state/b"])
        gen_states.append(cur)

    loss = 0
    for x, gx in zip(images[ctx:], gen_images[ctx - 1:]):
        loss = loss + F.mse_loss(gx, x)
    for s_, gs in zip(states[ctx:], gen_states[ctx - 1:]):
        loss = loss + F.mse_loss(gs, s_) * 1e-4
    loss = loss / float(T - ctx)
    return dict(loss=loss, P=P, gen_images=gen_images, masks=masks_out, gen_states=gen_states)


def loss_and_grads(params, batch, cfg, take_gt=None, dtype=torch.float64):
    out = forward(params, batch, cfg, take_gt, dtype)
    out["loss"].backward()
    out["grads"] = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).numpy() for k, v in out["P"].items()}
    return out
