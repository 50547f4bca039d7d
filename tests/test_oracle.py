"""CPU tests that pin the oracle (no GPU needed).

(i) oracle == independent torch autograd restatement (float64, fwd + every parameter gradient);
(ii) float64 finite differences for the hand-written operator backwards;
(iii) frozen golden vectors; (iv) the reference quirks of SURVEY Appendix B;
(v) bit-exact scheduled-sampling / context-frame index logic.
"""
import os
import sys

import numpy as np
import pytest

from oracle import model as M
from oracle import npgrad as G
from oracle import fused_ops as FO
import torch_restatement as TR

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _perturbed_params(cfg, seed=5, scale=0.1):
    p = M.init_params(cfg)
    rs = np.random.RandomState(seed)
    for k in sorted(p):
        if not k.endswith("/W"):
            p[k] = p[k] + scale * rs.standard_normal(p[k].shape).astype(p[k].dtype)
    return p


@pytest.mark.parametrize("mt,nm,k,oob", [
    ("CDNA", 10, 900.0, "zeros"), ("CDNA", 10, -1.0, "zeros"), ("CDNA", 3, 900.0, "zeros"),
    ("DNA", 1, 900.0, "zeros"), ("DNA", 1, -1.0, "zeros"),
    ("STP", 10, 900.0, "zeros"), ("STP", 10, -1.0, "border"), ("STP", 4, 900.0, "border"),
])
def test_oracle_matches_torch_restatement(mt, nm, k, oob):
    cfg = M.Config(mt, nm, schedsamp_k=k, height=16, width=24, dtype=np.float64, stp_oob=oob)
    p = _perturbed_params(cfg)
    batch = M.concat_examples(M.synthetic_sequences(3, 5, cfg))
    np.random.seed(99)
    take = []
    out = M.forward(p, batch, 6000, cfg, take_gt_log=take)
    G.backward(out["loss"])
    if k != -1.0:
        assert 0 < int(out["n_gt"]) < 3 and len(take) == 2      # mixing really happens
    tr = TR.loss_and_grads(p, batch, cfg, take_gt=take if k != -1.0 else None)
    assert abs(float(out["loss"].data) - float(tr["loss"].detach())) < 1e-12
    for a, b in zip(out["gen_images"], tr["gen_images"]):
        np.testing.assert_allclose(a.data, b.detach().numpy(), rtol=0, atol=1e-11)
    for tr_step, m in zip(out["trace"], tr["masks"]):
        np.testing.assert_allclose(tr_step["masks"].data, m.detach().numpy(), rtol=0, atol=1e-12)
    for key, v in out["P"].items():
        g = np.zeros_like(v.data) if v.grad is None else v.grad
        ref = tr["grads"][key]
        assert np.abs(g - ref).max() <= 1e-9 * (np.abs(ref).max() + 1e-12), key


def test_oracle_float32_close_to_float64():
    cfg64 = M.Config("CDNA", 10, height=16, width=16, dtype=np.float64)
    cfg32 = M.Config("CDNA", 10, height=16, width=16, dtype=np.float32)
    p = M.init_params(cfg32)
    batch = M.concat_examples(M.synthetic_sequences(2, 4, cfg32))
    o32 = M.loss_and_grads(p, batch, 0, cfg32)
    o64 = M.loss_and_grads(p, batch, 0, cfg64)
    assert o32["loss"].data.dtype == np.float32
    assert abs(float(o32["loss"].data) - float(o64["loss"].data)) < 1e-5 * abs(float(o64["loss"].data))
    for key in o32["grads"]:
        assert o32["grads"][key].dtype == np.float32, key
        den = np.abs(o64["grads"][key]).max() + 1e-12
        assert np.abs(o32["grads"][key] - o64["grads"][key]).max() / den < 2e-3, key


# ---------------------------------------------------------------------------- finite differences

def _fd_check(fn, arrays, wrt, eps=1e-6, n_probe=12, tol=1e-6):
    rs = np.random.RandomState(0)
    vs = [G.Var(a.copy()) for a in arrays]
    y = fn(*vs)
    seed = rs.standard_normal(y.data.shape)
    G.backward(y, seed=seed)
    for i in wrt:
        g = vs[i].grad
        flat = arrays[i].reshape(-1)
        for j in rs.choice(flat.size, min(n_probe, flat.size), replace=False):
            old = flat[j]
            flat[j] = old + eps
            yp = fn(*[G.Var(a.copy()) for a in arrays]).data
            flat[j] = old - eps
            ym = fn(*[G.Var(a.copy()) for a in arrays]).data
            flat[j] = old
            num = ((yp - ym) * seed).sum() / (2 * eps)
            assert abs(num - g.reshape(-1)[j]) <= tol * (1 + abs(num)), (i, j, num, g.reshape(-1)[j])


def test_fd_convolution_and_deconvolution():
    rs = np.random.RandomState(1)
    x = rs.standard_normal((2, 3, 8, 10))
    for (k, s, p) in ((5, 1, 2), (5, 2, 2), (3, 2, 1), (1, 1, 0)):
        W = rs.standard_normal((4, 3, k, k))
        b = rs.standard_normal(4)
        _fd_check(lambda a, w, c: G.convolution_2d(a, w, c, s, p), [x, W, b], (0, 1, 2))
    xd = rs.standard_normal((2, 3, 4, 5))
    Wd = rs.standard_normal((3, 4, 3, 3))
    bd = rs.standard_normal(4)
    _fd_check(lambda a, w, c: G.deconvolution_2d(a, w, c, 2, 1, (8, 10)), [xd, Wd, bd], (0, 1, 2))
    W1 = rs.standard_normal((3, 5, 1, 1))
    b1 = rs.standard_normal(5)
    _fd_check(lambda a, w, c: G.deconvolution_2d(a, w, c, 1, 0, (4, 5)), [xd, W1, b1], (0, 1, 2))


def test_fd_layernorm_depthwise_sampler_softmax():
    rs = np.random.RandomState(2)
    x = rs.standard_normal((3, 40))
    _fd_check(lambda a, g, b: G.layer_normalization(a, g, b), [x, rs.standard_normal(40), rs.standard_normal(40)],
              (0, 1, 2), tol=1e-5)
    xi = rs.standard_normal((3, 2, 7, 6))
    Wk = rs.standard_normal((4, 2, 5, 5))
    _fd_check(lambda a, w: G.depthwise_convolution_2d(a, w, 1, 2), [xi, Wk], (0, 1))
    _fd_check(lambda a: G.softmax(a), [rs.standard_normal((6, 11))], (0,))
    img = rs.standard_normal((2, 3, 6, 7))
    theta = np.tile(np.array([[1, 0, 0, 0, 1, 0.0]]), (2, 1)) + 0.3 * rs.standard_normal((2, 6))
    for oob in ("zeros", "border"):
        _fd_check(lambda a, th: G.spatial_transformer_sampler(
            a, G.spatial_transformer_grid(G.reshape(th, (2, 2, 3)), (6, 7)), oob), [img, theta], (0, 1), tol=1e-5)


def test_depthwise_matches_channel_order_A5():
    """out channel index is c*D + d (A.5) -- the CDNA un-flattening (ref:344) depends on it."""
    rs = np.random.RandomState(3)
    x = rs.standard_normal((1, 2, 6, 6))
    W = rs.standard_normal((3, 2, 5, 5))
    y = G.depthwise_convolution_2d(G.Var(x), G.Var(W), 1, 2).data
    import torch
    import torch.nn.functional as F
    ref = F.conv2d(torch.tensor(x), torch.tensor(W).transpose(0, 1).reshape(6, 1, 5, 5), padding=2, groups=2)
    np.testing.assert_allclose(y, ref.numpy(), atol=1e-12)


# ---------------------------------------------------------------------------- golden vectors

@pytest.mark.parametrize("name", ["cdna_sched", "cdna_feedself", "dna_sched", "stp_sched"])
def test_golden_vectors(name):
    sys.path.insert(0, GOLD)
    import make_golden
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    res = make_golden.run_case(name)
    assert int(res["n_gt"]) == int(gold["n_gt"])
    np.testing.assert_array_equal(res["take_gt"], gold["take_gt"])            # index work: bit-exact
    np.testing.assert_allclose(res["loss"], gold["loss"], rtol=2e-5)
    np.testing.assert_allclose(res["gen_last"], gold["gen_last"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(res["gen_state_last"], gold["gen_state_last"], rtol=0, atol=1e-5)
    for key in gold.files:
        if key.startswith("gl2/"):
            np.testing.assert_allclose(res[key], gold[key], rtol=2e-3, atol=1e-9, err_msg=key)


# ---------------------------------------------------------------------------- Appendix B quirks

def test_quirk_mask_softmax_groups_flat_nchw_elements():
    rs = np.random.RandomState(4)
    B, Mk, H, W = 2, 10, 8, 8
    prev = rs.rand(B, 3, H, W)
    a = rs.standard_normal((B, Mk + 1, H, W))
    r = FO.cdna_fused(prev, rs.standard_normal((B, 3, H, W)), a, rs.standard_normal((B, 25 * Mk)), Mk)
    flat = r["masks"].reshape(-1, Mk + 1)
    np.testing.assert_allclose(flat.sum(1), 1.0, atol=1e-12)                   # flat groups sum to 1
    per_pixel = r["masks"].sum(1)
    assert np.abs(per_pixel - 1.0).max() > 0.05                                # ... channels of a pixel do not
    z = np.maximum(a, 0).reshape(-1, Mk + 1)
    e = np.exp(z - z.max(1, keepdims=True))
    np.testing.assert_allclose(flat, e / e.sum(1, keepdims=True), atol=1e-12)


def test_quirk_cdna_last_kernel_unused_zero_grad():
    rs = np.random.RandomState(5)
    B, Mk, H, W = 2, 10, 8, 8
    r = FO.cdna_fused(rs.rand(B, 3, H, W), rs.standard_normal((B, 3, H, W)),
                      rs.standard_normal((B, Mk + 1, H, W)), rs.rand(B, 25 * Mk) + 0.1, Mk)
    g = r["bwd"](rs.standard_normal((B, 3, H, W)))
    gk = g["kern_raw"].reshape(B, Mk, 25)
    assert np.all(gk[:, Mk - 1] == 0)                                          # B.3: exactly zero
    assert np.abs(gk[:, :Mk - 1]).max() > 0


def test_quirk_dna_taps_truncated_and_detached():
    rs = np.random.RandomState(6)
    B, H, W = 1, 8, 8
    prev = rs.rand(B, 3, H, W)
    e = np.full((B, 25, H, W), -1.0)
    e[:, 24] = 1.0                    # only tap (xk=4, yk=4): out[i,j] = prev[i+2, j+2] inside the truncated window
    a = np.zeros((B, 2, H, W))
    r = FO.dna_fused(prev, e, a)
    t = r["transformed"][0, 0]
    np.testing.assert_allclose(t[:H - 4, :W - 4], prev[0, 0, 2:H - 2, 2:W - 2], atol=1e-9)
    assert np.abs(t[H - 4:, :]).max() < 1e-9 and np.abs(t[:, W - 4:]).max() < 1e-9   # two extra zero rows/cols
    g_out = rs.standard_normal((B, 3, H, W))
    g = r["bwd"](g_out)
    np.testing.assert_allclose(g["prev"], r["masks"][:, 0:1] * g_out, atol=1e-12)  # only via prev*mask0


def test_quirk_stp_transformers_identical():
    rs = np.random.RandomState(7)
    cfg = M.Config("STP", 4, height=16, width=16, dtype=np.float64)
    p = M.init_params(cfg)
    batch = M.concat_examples(M.synthetic_sequences(2, 3, cfg))
    out = M.forward(p, batch, 0, cfg)
    tl = out["trace"][0]["transformed"]
    assert len(tl) == 4
    np.testing.assert_array_equal(tl[1].data, tl[2].data)
    np.testing.assert_array_equal(tl[1].data, tl[3].data)


# ---------------------------------------------------------------------------- scheduled sampling, indexing

def test_num_ground_truth_schedule_known_values():
    # b32, k=900
    assert [int(M.num_ground_truth(32, 900.0, it)) for it in (0, 3000, 6000, 10000)] == [32, 31, 17, 0]
    assert M.num_ground_truth(32, 900.0, 0).dtype == np.int32


@pytest.mark.parametrize("n_gt", [0, 1, 17, 31, 32])
def test_scheduled_sample_select_form_is_bit_exact(n_gt):
    rs = np.random.RandomState(8)
    gt = rs.rand(32, 3, 4, 4).astype(np.float32)
    gen = rs.rand(32, 3, 4, 4).astype(np.float32)
    np.random.seed(1234)
    lit = M.scheduled_sample(gt, gen, 32, n_gt)
    np.random.seed(1234)
    take = M.scheduled_sample_order(32, n_gt)
    sel = np.where(take[:, None, None, None], gt, gen)
    assert take.sum() == n_gt
    np.testing.assert_array_equal(lit, sel)


def test_context_frame_indexing_and_loss_terms():
    cfg = M.Config("CDNA", 10, schedsamp_k=900.0, height=16, width=16)
    p = M.init_params(cfg)
    T = 6
    batch = M.concat_examples(M.synthetic_sequences(4, T, cfg))
    np.random.seed(0)
    take = []
    out = M.forward(p, batch, 6000, cfg, take_gt_log=take)
    assert len(out["gen_images"]) == T - 1
    assert len(take) == T - 1 - cfg.context_frames          # one shuffle per step with len(gen) > ctx-1
    assert len(out["recon_costs"]) == T - cfg.context_frames
    # first two steps are fed ground truth
    np.testing.assert_array_equal(out["trace"][0]["prev_image"].data, batch[0][0])
    np.testing.assert_array_equal(out["trace"][1]["prev_image"].data, batch[0][1])
    sel = np.where(take[0][:, None, None, None], batch[0][2], out["gen_images"][1].data)
    np.testing.assert_array_equal(out["trace"][2]["prev_image"].data, sel)


def test_concat_examples_layout():
    cfg = M.Config(height=16, width=24)
    seqs = M.synthetic_sequences(3, 4, cfg)
    img, act, sta = M.concat_examples(seqs)
    assert img.shape == (4, 3, 3, 16, 24) and act.shape == (4, 3, 5) and sta.shape == (4, 3, 5)
    np.testing.assert_array_equal(img[2, 1, :, 5, 7], seqs[1][0][2, 5, 7, :])
    np.testing.assert_array_equal(act[3, 2], seqs[2][1][3])


def test_adam_rule_matches_closed_form():
    adam = M.Adam(alpha=1e-3)
    p = {"w": np.array([1.0, -2.0], np.float32)}
    g = {"w": np.array([0.5, -0.25], np.float32)}
    adam.update(p, g)
    # t=1: m=(1-b1)g, v=(1-b2)g^2, lr = a*sqrt(1-b2)/(1-b1)
    m = 0.1 * g["w"]
    v = 0.001 * g["w"] ** 2
    lr = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(p["w"], np.array([1.0, -2.0]) - lr * m / (np.sqrt(v) + 1e-8), rtol=1e-6)
