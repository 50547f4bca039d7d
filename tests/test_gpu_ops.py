"""GPU parity tests, op by op, through the C-ABI (ctypes -> libpivp.so) against the NumPy oracle.

Tolerances (fp32 path): forward values 1e-4 relative to the tensor's max-abs (north_star: "within 1e-4 relative in
fp32"); gradients 1e-3 relative to the tensor's max-abs (stated per test); index / select work bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import npgrad as G
from oracle import fused_ops as FO
from oracle import model as OM

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 1e-3


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a.astype(np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


@pytest.fixture(scope="module")
def pk():
    import pivp_b200
    pivp_b200.lib()
    return pivp_b200


# ---------------------------------------------------------------------------------------------- convolutions
@pytest.mark.parametrize("B,C,H,W,O,k,s,p", [
    (2, 3, 16, 24, 32, 5, 2, 2),      # enc0 (generic implicit-GEMM kernels: output not tileable by 8 x 16)
    (2, 3, 32, 64, 32, 5, 2, 2),      # enc0 on the direct image kernels (conv_image.cu)
    (3, 3, 64, 64, 32, 5, 2, 2),
    (2, 32, 16, 16, 32, 3, 2, 1),     # enc1
    (3, 74, 8, 8, 64, 1, 1, 0),       # enc3 (74 input channels: not a multiple of 4)
    (2, 64, 16, 16, 128, 5, 1, 2),    # a ConvLSTM gate conv
    (1, 96, 8, 12, 256, 5, 1, 2),
    (2, 5, 7, 9, 6, 3, 1, 1),         # ragged everything
])
def test_conv2d_fwd_bwd(pk, B, C, H, W, O, k, s, p):
    rs = np.random.RandomState(0)
    x = rs.standard_normal((B, C, H, W)).astype(np.float32)
    Wt = (rs.standard_normal((O, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    b = rs.standard_normal(O).astype(np.float32)
    vx, vw, vb = G.Var(x.astype(np.float64)), G.Var(Wt.astype(np.float64)), G.Var(b.astype(np.float64))
    y = G.convolution_2d(vx, vw, vb, s, p)
    gy = rs.standard_normal(y.data.shape).astype(np.float32)
    G.backward(y, seed=gy.astype(np.float64))
    f = pk.functions.Convolution2DFunction(s, p)
    (yg,) = f.forward((cu(x), cu(Wt), cu(b)))
    dx, dw, db = f.backward((cu(x), cu(Wt), cu(b)), (cu(gy),))
    assert rel(yg, y.data) < FWD_TOL
    assert rel(dx, vx.grad) < GRAD_TOL and rel(dw, vw.grad) < GRAD_TOL and rel(db, vb.grad) < GRAD_TOL


@pytest.mark.parametrize("B,Ci,ih,iw,Co,k,s,p", [
    (2, 128, 4, 4, 128, 3, 2, 1),     # enc4
    (2, 96, 8, 6, 96, 3, 2, 1),       # enc5
    (1, 64, 16, 16, 64, 3, 2, 1),     # enc6
    (2, 64, 8, 8, 11, 1, 1, 0),       # masks head (1x1 deconv)
])
def test_deconv2d_fwd_bwd(pk, B, Ci, ih, iw, Co, k, s, p):
    rs = np.random.RandomState(1)
    oh, ow = (ih * 2, iw * 2) if s == 2 else (ih, iw)
    x = rs.standard_normal((B, Ci, ih, iw)).astype(np.float32)
    Wt = (rs.standard_normal((Ci, Co, k, k)) / np.sqrt(Co * k * k)).astype(np.float32)
    b = rs.standard_normal(Co).astype(np.float32)
    vx, vw, vb = G.Var(x.astype(np.float64)), G.Var(Wt.astype(np.float64)), G.Var(b.astype(np.float64))
    y = G.deconvolution_2d(vx, vw, vb, s, p, (oh, ow))
    gy = rs.standard_normal(y.data.shape).astype(np.float32)
    G.backward(y, seed=gy.astype(np.float64))
    f = pk.functions.Deconvolution2DFunction(s, p, (oh, ow))
    (yg,) = f.forward((cu(x), cu(Wt), cu(b)))
    dx, dw, db = f.backward((cu(x), cu(Wt), cu(b)), (cu(gy),))
    assert rel(yg, y.data) < FWD_TOL
    assert rel(dx, vx.grad) < GRAD_TOL and rel(dw, vw.grad) < GRAD_TOL and rel(db, vb.grad) < GRAD_TOL


def test_conv2d_strided_views_and_errors(pk):
    """Channel-slice views (the free F.concat) and the error contract of the C-ABI."""
    rs = np.random.RandomState(2)
    B, H, W = 2, 8, 8
    a = rs.standard_normal((B, 5, H, W)).astype(np.float32)
    h = rs.standard_normal((B, 3, H, W)).astype(np.float32)
    Wt = rs.standard_normal((4, 8, 3, 3)).astype(np.float32)
    y = G.convolution_2d(G.concat((G.Var(a), G.Var(h))), G.Var(Wt), G.Var(np.zeros(4, np.float32)), 1, 1).data
    L = pk.lib()
    s = torch.cuda.current_stream().cuda_stream
    buf = torch.zeros(B * H * W, 8 + 2, device="cuda")           # row stride 10, slice at offset 1
    L.call("pivp_nchw_to_nhwc", cu(a).data_ptr(), buf.data_ptr() + 4, 10, 0, B, 5, H * W, s)
    L.call("pivp_nchw_to_nhwc", cu(h).data_ptr(), buf.data_ptr() + 4, 10, 5, B, 3, H * W, s)
    wi = cu(Wt.transpose(0, 2, 3, 1))
    out = torch.full((B * H * W, 6), 7.0, device="cuda")
    L.call("pivp_conv2d_fwd", buf.data_ptr(), 10, 1, B, H, W, 8, wi.data_ptr(), 0, 4, 3, 3, 1, 1, out.data_ptr(), 6, 2, H, W, 0, 0, s)
    got = out[:, 2:6].reshape(B, H, W, 4).permute(0, 3, 1, 2)
    assert rel(got, y) < FWD_TOL
    assert torch.all(out[:, :2] == 7.0)                                # outside the slice untouched
    with pytest.raises(pk.PivpError):
        L.call("pivp_conv2d_fwd", buf.data_ptr(), 10, 1, B, H, W, 8, wi.data_ptr(), 0, 4, 3, 3, 1, 1, out.data_ptr(), 6, 2, H + 1, W, 0, 0, s)
    with pytest.raises(pk.PivpError):
        L.call("pivp_conv2d_fwd", 0, 10, 1, B, H, W, 8, wi.data_ptr(), 0, 4, 3, 3, 1, 1, out.data_ptr(), 6, 2, H, W, 0, 0, s)
    assert b"conv" in L.cdll.pivp_last_error()


# ---------------------------------------------------------------------------------------------- LayerNorm, ConvLSTM
@pytest.mark.parametrize("B,C,H,W,relu", [(3, 32, 8, 8, False), (2, 64, 16, 16, True), (2, 128, 2, 3, False), (1, 64, 64, 64, True),
                                           # one-launch cluster kernels both ways: n = 4096 CL, CL = 8 / 2 / 1
                                           (2, 32, 32, 32, False), (3, 128, 8, 8, True), (2, 16, 16, 16, False)])
def test_layernorm_fwd_bwd(pk, B, C, H, W, relu):
    rs = np.random.RandomState(3)
    n = C * H * W
    x = (rs.standard_normal((B, C, H, W)) * 2 + 0.5).astype(np.float32)
    ga = (1 + 0.2 * rs.standard_normal(n)).astype(np.float32)
    be = (0.2 * rs.standard_normal(n)).astype(np.float32)
    vx, vg, vb = G.Var(x.astype(np.float64)), G.Var(ga.astype(np.float64)), G.Var(be.astype(np.float64))
    y = G.reshape(G.layer_normalization(G.reshape(vx, (B, -1)), vg, vb), x.shape)
    if relu:
        y = G.relu(y)
    gy = rs.standard_normal(x.shape).astype(np.float32)
    G.backward(y, seed=gy.astype(np.float64))
    f = pk.functions.LayerNormalizationFunction(relu)
    (yg,) = f.forward((cu(x), cu(ga), cu(be)))
    dx, dg, db = f.backward((cu(x), cu(ga), cu(be)), (cu(gy),))
    assert rel(yg, y.data) < FWD_TOL
    assert rel(dx, vx.grad) < GRAD_TOL and rel(dg, vg.grad) < GRAD_TOL and rel(db, vb.grad) < GRAD_TOL


def test_basic_conv_lstm_cell_two_steps(pk):
    rs = np.random.RandomState(4)
    B, Cin, C, H, W = 2, 32, 64, 8, 8
    cell = pk.BasicConvLSTMCell(C)
    xs = [rs.standard_normal((B, Cin, H, W)).astype(np.float32) for _ in range(2)]
    Wt = (rs.standard_normal((4 * C, Cin + C, 5, 5)) / np.sqrt((Cin + C) * 25)).astype(np.float32)
    bt = (0.1 * rs.standard_normal(4 * C)).astype(np.float32)
    cell.W, cell.b = cu(Wt), cu(bt)
    P = {"l/conv/W": G.Var(Wt.astype(np.float64)), "l/conv/b": G.Var(bt.astype(np.float64))}
    st = {}
    for x in xs:
        h_ref = OM.conv_lstm(P, "l", G.Var(x.astype(np.float64)), st, C)
        h = cell(cu(x))
    assert rel(h, h_ref.data) < FWD_TOL
    assert rel(cell.c, st["l"][0].data) < FWD_TOL
    cell.reset_state()
    assert cell.c is None and cell.h is None


# ---------------------------------------------------------------------------------------------- fused transforms
def _fused_inputs(rs, B, H, W, M, ne):
    prev = rs.rand(B, 3, H, W).astype(np.float32)
    e = rs.standard_normal((B, ne, H, W)).astype(np.float32)
    a = (2 * rs.standard_normal((B, M + 1, H, W))).astype(np.float32)
    g = rs.standard_normal((B, 3, H, W)).astype(np.float32)
    return prev, e, a, g


@pytest.mark.parametrize("B,H,W,M", [(2, 64, 64, 10), (3, 16, 24, 3), (1, 8, 8, 1), (2, 24, 40, 7)])
def test_cdna_fused_fwd_bwd(pk, B, H, W, M):
    rs = np.random.RandomState(5)
    prev, e, a, g = _fused_inputs(rs, B, H, W, M, 3)
    k = rs.standard_normal((B, 25 * M)).astype(np.float32)       # about half the taps are clamped to RELU_SHIFT
    ref = FO.cdna_fused(*(v.astype(np.float64) for v in (prev, e, a, k)), M)
    gr = ref["bwd"](g.astype(np.float64))
    f = pk.functions.CDNACompositeFunction(M)
    ins = tuple(cu(v) for v in (prev, e, a, k))
    (out,) = f.forward(ins)
    dp, de, da, dk = f.backward(ins, (cu(g),))
    assert rel(out, ref["out"]) < FWD_TOL
    assert rel(de, gr["enc7_pre"]) < GRAD_TOL and rel(da, gr["mask_pre"]) < GRAD_TOL
    assert rel(dk, gr["kern_raw"]) < GRAD_TOL and rel(dp, gr["prev"]) < GRAD_TOL
    dk3 = dk.reshape(B, M, 25)
    assert torch.all(dk3[:, M - 1] == 0)                         # B.3: the last kernel gets exactly zero gradient


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (3, 16, 24), (1, 8, 8)])
def test_dna_fused_fwd_bwd(pk, B, H, W):
    rs = np.random.RandomState(6)
    prev, e, a, g = _fused_inputs(rs, B, H, W, 1, 25)
    ref = FO.dna_fused(*(v.astype(np.float64) for v in (prev, e, a)))
    gr = ref["bwd"](g.astype(np.float64))
    f = pk.functions.DNACompositeFunction()
    ins = tuple(cu(v) for v in (prev, e, a))
    (out,) = f.forward(ins)
    dp, de, da = f.backward(ins, (cu(g),))
    assert rel(out, ref["out"]) < FWD_TOL
    assert rel(de, gr["enc7_pre"]) < GRAD_TOL and rel(da, gr["mask_pre"]) < GRAD_TOL and rel(dp, gr["prev"]) < GRAD_TOL


@pytest.mark.parametrize("B,H,W,M,oob", [(2, 64, 64, 10, "zeros"), (2, 64, 64, 10, "border"), (3, 16, 24, 4, "zeros"),
                                         (2, 128, 128, 10, "zeros")])
def test_stp_fused_fwd_bwd(pk, B, H, W, M, oob):
    rs = np.random.RandomState(7)
    prev, e, a, g = _fused_inputs(rs, B, H, W, M, 3)
    th_raw = (0.4 * rs.standard_normal((B, 6))).astype(np.float32)          # a good share of samples fall outside
    ident = np.array([1, 0, 0, 0, 1, 0], np.float64)
    ref = FO.stp_fused(prev.astype(np.float64), e.astype(np.float64), a.astype(np.float64), th_raw.astype(np.float64) + ident, M, oob)
    gr = ref["bwd"](g.astype(np.float64))
    f = pk.functions.STPCompositeFunction(M, oob)
    ins = tuple(cu(v) for v in (prev, e, a, th_raw))
    (out,) = f.forward(ins)
    dp, de, da, dth = f.backward(ins, (cu(g),))
    # the bilinear sampler amplifies the fp32 rounding of the grid coordinates by the local image slope
    assert rel(out, ref["out"]) < 5e-4
    assert rel(de, gr["enc7_pre"]) < GRAD_TOL and rel(da, gr["mask_pre"]) < 2e-3
    assert rel(dth, gr["theta"]) < 5e-3 and rel(dp, gr["prev"]) < 2e-3


def test_cdna_fused_full_size_b32_against_per_sample_oracle(pk):
    """BASELINE size (b32, 64x64, 10 masks): every op is per-sample, so check samples against the oracle one by one."""
    rs = np.random.RandomState(8)
    B, H, W, M = 32, 64, 64, 10
    prev, e, a, g = _fused_inputs(rs, B, H, W, M, 3)
    k = rs.standard_normal((B, 25 * M)).astype(np.float32)
    f = pk.functions.CDNACompositeFunction(M)
    ins = tuple(cu(v) for v in (prev, e, a, k))
    (out,) = f.forward(ins)
    dp, de, da, dk = f.backward(ins, (cu(g),))
    for b in (0, 13, 31):
        sl = slice(b, b + 1)
        ref = FO.cdna_fused(prev[sl].astype(np.float64), e[sl].astype(np.float64), a[sl].astype(np.float64), k[sl].astype(np.float64), M)
        gr = ref["bwd"](g[sl].astype(np.float64))
        assert rel(out[sl], ref["out"]) < FWD_TOL
        assert rel(da[sl], gr["mask_pre"]) < GRAD_TOL and rel(dk[sl], gr["kern_raw"]) < GRAD_TOL
    # size-independent properties: samples are independent (batch reversal commutes, bit-exact) and backward is linear in g
    rev = tuple(torch.flip(v, dims=[0]).contiguous() for v in ins)
    (out_r,) = f.forward(rev)
    assert torch.equal(torch.flip(out_r, dims=[0]), out)
    dp2, de2, da2, dk2 = f.backward(ins, (cu(2 * g),))
    assert rel(da2, 2 * da.cpu().numpy().astype(np.float64)) < 1e-5 and rel(dk2, 2 * dk.cpu().numpy().astype(np.float64)) < 1e-4


# ---------------------------------------------------------------------------------------------- small ops
def test_sched_select_bit_exact(pk):
    rs = np.random.RandomState(9)
    B = 32
    gt, gen = rs.rand(B, 3, 8, 8).astype(np.float32), rs.rand(B, 3, 8, 8).astype(np.float32)
    for n_gt in (0, 1, 17, 31, 32):
        np.random.seed(5)
        ref = OM.scheduled_sample(gt, gen, B, n_gt)
        np.random.seed(5)
        out = pk.scheduled_sample(cu(gt), cu(gen), B, n_gt)
        assert np.array_equal(out.cpu().numpy(), ref)
        # the variant the training step uses: the same select plus its NHWC rows (what enc0 reads)
        take = torch.from_numpy((ref.reshape(B, -1) == gt.reshape(B, -1)).all(1).astype(np.int32)).cuda()
        o1, o2 = torch.empty(B, 3, 8, 8, device="cuda"), torch.empty(B, 8, 8, 3, device="cuda")
        gt_d, gen_d = cu(gt), cu(gen)
        pk.lib().call("pivp_sched_select_nhwc", gt_d.data_ptr(), gen_d.data_ptr(), take.data_ptr(), o1.data_ptr(), o2.data_ptr(), B, 3, 64,
                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(o1.cpu().numpy(), ref) and np.array_equal(o2.permute(0, 3, 1, 2).cpu().numpy(), ref)


def test_adam_matches_chainer_rule(pk):
    rs = np.random.RandomState(10)
    n = 100003
    p0 = rs.standard_normal(n).astype(np.float32)
    oa = OM.Adam()
    params = {"w": p0.copy()}
    p = cu(p0)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        g = (rs.standard_normal(n) * 10.0 ** rs.randint(-6, 1, n)).astype(np.float32)
        oa.update(params, {"w": g})
        pk.lib().call("pivp_adam_step", p.data_ptr(), cu(g).data_ptr(), m.data_ptr(), v.data_ptr(), n, step.data_ptr(),
                      1e-3, 0.9, 0.999, 1e-8, 1.0, s)
    assert int(step.item()) == 3
    assert rel(p, params["w"]) < 1e-6


@pytest.mark.parametrize("B,H,W,Ne,M1", [(2, 16, 24, 3, 11), (3, 8, 8, 25, 2), (1, 64, 64, 3, 11)])
def test_heads_fused_fwd_bwd(pk, B, H, W, Ne, M1):
    """The two 1x1 head deconvolutions (enc7 + masks) fused: NCHW planes out / in, against the oracle's Deconvolution2D."""
    rs = np.random.RandomState(21)
    L = pk.lib()
    s = torch.cuda.current_stream().cuda_stream
    NH, HW, M = Ne + M1, H * W, B * H * W
    x = rs.standard_normal((B, 64, H, W)).astype(np.float32)
    Wa, Wb = (rs.standard_normal((64, Ne, 1, 1)) / 8).astype(np.float32), (rs.standard_normal((64, M1, 1, 1)) / 8).astype(np.float32)
    ba, bb = rs.standard_normal(Ne).astype(np.float32), rs.standard_normal(M1).astype(np.float32)
    vx = G.Var(x.astype(np.float64))
    va, vb, vba, vbb = (G.Var(v.astype(np.float64)) for v in (Wa, Wb, ba, bb))
    ya = G.deconvolution_2d(vx, va, vba, 1, 0, (H, W))
    yb = G.deconvolution_2d(vx, vb, vbb, 1, 0, (H, W))
    ga, gb = rs.standard_normal(ya.data.shape).astype(np.float32), rs.standard_normal(yb.data.shape).astype(np.float32)
    G.backward(G.sum_(ya * G.Var(ga.astype(np.float64)), (0, 1, 2, 3)) + G.sum_(yb * G.Var(gb.astype(np.float64)), (0, 1, 2, 3)))
    # internal layouts: x NHWC rows inside a wider buffer (stride 72, offset 4), W [NH][64], b [NH]
    xb = torch.zeros(M, 72, device="cuda")
    xb[:, 4:68] = cu(x.transpose(0, 2, 3, 1).reshape(M, 64))
    Wi = cu(np.concatenate([Wa[:, :, 0, 0].T, Wb[:, :, 0, 0].T], 0))
    bi = cu(np.concatenate([ba, bb]))
    oa, ob = torch.empty(B, Ne, H, W, device="cuda"), torch.empty(B, M1, H, W, device="cuda")
    L.call("pivp_heads_fwd", xb.data_ptr(), 72, 4, Wi.data_ptr(), bi.data_ptr(), oa.data_ptr(), Ne, ob.data_ptr(), NH, B, HW, s)
    assert rel(oa, ya.data) < FWD_TOL and rel(ob, yb.data) < FWD_TOL
    dx = torch.full((M, 72), 3.0, device="cuda")
    dW, db = torch.full((NH, 64), 0.5, device="cuda"), torch.full((NH,), 0.25, device="cuda")      # accumulated into
    gad, gbd = cu(ga), cu(gb)                                   # keep the device buffers alive across the call
    L.call("pivp_heads_bwd", xb.data_ptr(), 72, 4, Wi.data_ptr(), gad.data_ptr(), Ne, gbd.data_ptr(), NH, dx.data_ptr(), 72, 4,
           dW.data_ptr(), db.data_ptr(), B, HW, s)
    assert rel(dx[:, 4:68].reshape(B, H, W, 64).permute(0, 3, 1, 2), vx.grad) < GRAD_TOL
    assert torch.all(dx[:, :4] == 3.0) and torch.all(dx[:, 68:] == 3.0)                             # outside the view untouched
    dW_ref = np.concatenate([va.grad[:, :, 0, 0].T, vb.grad[:, :, 0, 0].T], 0)
    assert rel(dW - 0.5, dW_ref) < GRAD_TOL and rel(db - 0.25, np.concatenate([vba.grad, vbb.grad])) < GRAD_TOL


@pytest.mark.parametrize("M,C,H,W,s2d,mask,two", [(2 * 16 * 16, 96, 16, 16, 1, 1, 0), (3 * 8 * 8, 64, 8, 8, 1, 0, 0), (500, 32, 0, 0, 0, 1, 1)])
def test_grad_handover_matches_separate_kernels(pk, M, C, H, W, s2d, mask, two):
    """relu_bwd + colsum + cast_bf16 in one launch: bit-identical fp32 / bf16 outputs, column sums to fp32 rounding."""
    L = pk.lib()
    s = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    out = torch.randn(M, C + 8, device="cuda", generator=gen)
    ga, gb = torch.randn(M, C + 4, device="cuda", generator=gen), torch.randn(M, C, device="cuda", generator=gen)
    cb = (C + 63) // 64 * 64
    rows = M // 4 if s2d else M
    bcs = 4 * cb if s2d else cb
    d1, d2 = torch.zeros(M, C, device="cuda"), torch.zeros(M, C, device="cuda")
    b1, b2 = torch.zeros(rows, bcs, dtype=torch.bfloat16, device="cuda"), torch.zeros(rows, bcs, dtype=torch.bfloat16, device="cuda")
    db1, db2 = torch.full((C,), 0.5, device="cuda"), torch.full((C,), 0.5, device="cuda")
    L.call("pivp_grad_handover", out.data_ptr() if mask else 0, C + 8, 4, ga.data_ptr(), C + 4, 0, gb.data_ptr() if two else 0, C, 0,
           d1.data_ptr(), C, 0, b1.data_ptr(), bcs, 0, H, W, s2d, cb, db1.data_ptr(), M, C, s)
    if mask:
        L.call("pivp_relu_bwd", out.data_ptr(), C + 8, 4, ga.data_ptr(), C + 4, 0, gb.data_ptr() if two else 0, C, 0, d2.data_ptr(), C, 0, M, C, s)
    else:
        d2.copy_(ga[:, :C] + (gb if two else 0))
    L.call("pivp_colsum", d2.data_ptr(), C, 0, M, C, db2.data_ptr(), s)
    L.call("pivp_cast_bf16", d2.data_ptr(), C, 0, b2.data_ptr(), bcs, 0, M, C, H, W, s2d, cb, s)
    torch.cuda.synchronize()
    assert torch.equal(d1, d2) and torch.equal(b1, b2)
    assert rel(db1, db2.double().cpu().numpy()) < 1e-5


@pytest.mark.parametrize("B,K,N,relu", [(32, 8192, 250, 0), (7, 2048, 100, 1), (3, 1024, 6, 0)])
def test_linear_wide_paths_match_oracle(pk, B, K, N, relu):
    """The split-K forward and the wide backward kernels (weight matrix streamed once) against the oracle's Linear."""
    rs = np.random.RandomState(12)
    L = pk.lib()
    s = torch.cuda.current_stream().cuda_stream
    x, W, b = rs.standard_normal((B, K)).astype(np.float32), (rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32), rs.standard_normal(N).astype(np.float32)
    vx, vw, vb = G.Var(x.astype(np.float64)), G.Var(W.astype(np.float64)), G.Var(b.astype(np.float64))
    y = G.linear(vx, vw, vb)
    if relu:
        y = G.relu(y)
    gy = rs.standard_normal((B, N)).astype(np.float32)
    G.backward(y, seed=gy.astype(np.float64))
    xg, Wg, bg = cu(x), cu(W), cu(b)
    yg = torch.empty(B, N, device="cuda")
    nb = L.query("pivp_linear_fwd_workspace_bytes", B, K, N)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    L.call("pivp_linear_fwd_splitk", xg.data_ptr(), K, Wg.data_ptr(), bg.data_ptr(), yg.data_ptr(), B, K, N, relu, ws.data_ptr(), nb, s)
    assert rel(yg, y.data) < FWD_TOL
    gyg = cu(gy * (y.data > 0) if relu else gy)
    dx = torch.full((B, K), 7.0, device="cuda")                         # overwritten (accumulate_dx = 0)
    dW0, db0 = rs.standard_normal((N, K)).astype(np.float32), rs.standard_normal(N).astype(np.float32)
    dW, db = cu(dW0), cu(db0)                                           # accumulated into
    L.call("pivp_linear_bwd", gyg.data_ptr(), xg.data_ptr(), K, Wg.data_ptr(), dx.data_ptr(), K, 0, dW.data_ptr(), db.data_ptr(), B, K, N, s)
    assert rel(dx, vx.grad) < GRAD_TOL and rel(dW, dW0 + vw.grad) < GRAD_TOL and rel(db, db0 + vb.grad) < GRAD_TOL
    dx2 = dx.clone()
    L.call("pivp_linear_bwd", gyg.data_ptr(), xg.data_ptr(), K, Wg.data_ptr(), dx2.data_ptr(), K, 1, dW.data_ptr(), db.data_ptr(), B, K, N, s)
    assert rel(dx2, 2 * vx.grad) < GRAD_TOL


def test_linear_mse_state(pk):
    rs = np.random.RandomState(11)
    L = pk.lib()
    s = torch.cuda.current_stream().cuda_stream
    B, K, N = 5, 300, 250
    x, W, b = rs.standard_normal((B, K)).astype(np.float32), rs.standard_normal((N, K)).astype(np.float32), rs.standard_normal(N).astype(np.float32)
    vx, vw, vb = G.Var(x.astype(np.float64)), G.Var(W.astype(np.float64)), G.Var(b.astype(np.float64))
    y = G.linear(vx, vw, vb)
    gy = rs.standard_normal((B, N)).astype(np.float32)
    G.backward(y, seed=gy.astype(np.float64))
    xg, Wg, bg, gyg = cu(x), cu(W), cu(b), cu(gy)
    yg = torch.empty(B, N, device="cuda")
    L.call("pivp_linear_fwd", xg.data_ptr(), K, Wg.data_ptr(), bg.data_ptr(), yg.data_ptr(), B, K, N, 0, s)
    dx, dW, db = torch.empty(B, K, device="cuda"), torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    L.call("pivp_linear_bwd", gyg.data_ptr(), xg.data_ptr(), K, Wg.data_ptr(), dx.data_ptr(), K, 0, dW.data_ptr(), db.data_ptr(), B, K, N, s)
    assert rel(yg, y.data) < FWD_TOL and rel(dx, vx.grad) < GRAD_TOL and rel(dW, vw.grad) < GRAD_TOL and rel(db, vb.grad) < GRAD_TOL
    # mse
    a_, b_ = rs.rand(7, 3, 9, 5).astype(np.float32), rs.rand(7, 3, 9, 5).astype(np.float32)
    slot, dg = torch.zeros(1, device="cuda"), torch.empty(a_.size, device="cuda")
    ag, bgt = cu(a_), cu(b_)                       # keep the device buffers alive across the call
    L.call("pivp_mse", ag.data_ptr(), bgt.data_ptr(), a_.size, 2.0 / a_.size, dg.data_ptr(), slot.data_ptr(), s)
    assert abs(float(slot.item()) / a_.size - ((a_ - b_) ** 2).mean()) < 1e-6
    assert rel(dg.reshape(a_.shape), 2 * (a_.astype(np.float64) - b_) / a_.size) < 1e-5
    assert abs(pk.peak_signal_to_noise_ratio(cu(b_), cu(a_)) - 10 * np.log10(1 / ((a_ - b_) ** 2).mean())) < 1e-3
