"""Where does the bf16 gradient error come from?  A/B on the fp32 (SIMT) engine with ONE bf16 rounding switched on at a time,
against the float64 oracle -- the attribution VERDICT r1 asked for (weak #3):

  gates : the activated ConvLSTM gates saved for backward are rounded to bf16 (what the tensor-core path stores in dg_bf16)
  dG    : the gate pre-activation gradients are rounded to bf16 before the input- and weight-gradient convolutions
  xh    : the concatenated ConvLSTM input [x | h_{t-1}] is rounded to bf16 as GEMM operand (forward conv and weight gradient);
          the fp32 h that the LayerNorm and the recurrence see stays fp32, as on the tensor-core path
  W     : every weight the tensor-core path casts to bf16 (ConvLSTM, enc1/2, enc4/5/6) is rounded once
  all   : the four together;   bf16 : the real tensor-core engine (adds bf16 deconvolution activations and tcgen05 accumulation order)

    python scripts/diag_bf16_ab.py [B=2] [T=4] [CDNA|DNA|STP]        # 64x64, scheduled sampling at iteration 6000, perturbed parameters
torch is used for the roundings only (diagnostics, not the product path)."""
import os, sys
os.environ.setdefault("PIVP_BRANCHES", "")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk
from pivp_b200 import engine as E
from oracle import model as OM, npgrad as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
MT = sys.argv[3] if len(sys.argv) > 3 else "CDNA"
NM = 1 if MT == "DNA" else 10
H = W = 64
cfg = OM.Config(MT, NM, schedsamp_k=900.0, height=H, width=W, dtype=np.float64)
params = OM.init_params(cfg)
rs = np.random.RandomState(7)
for key in sorted(params):
    if not key.endswith("/W"):
        params[key] = params[key] + 0.05 * rs.standard_normal(params[key].shape)
batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
np.random.seed(99); ref = OM.forward(params, batch, 6000, cfg); G.backward(ref["loss"])
rb = lambda t: t.copy_(t.bfloat16().float())            # round a tensor to bf16 precision in place
View, _ptr, LI, LS, LV = E.View, E._ptr, E.LSTM_IN, E.LSTM_SIZES, E.LSTM_LEVEL


def patched(eng, sw):
    xr = {}

    def lstm_fwd(li, t, Bn):
        ws = eng.ws
        cin, C, lv = LI[li], LS[li], LV[li]
        h, w = eng.H // lv, eng.W // lv
        name = "lstm%d/conv" % (li + 1)
        xh, Gt = ws["xh"][li][t], ws["G"][li][t]
        src = xh
        if "xh" in sw:
            src = xr[(li, t)] = xh.bfloat16().float()
        eng._conv_fwd(View(src, cin + C, 0, cin + C), Bn, h, w, eng.p[name + "/W"], eng.p[name + "/b"], 4 * C, 5, 1, 2, View(Gt, 4 * C, 0, 4 * C))
        eng.L.call("pivp_lstm_gates_fwd", _ptr(Gt), _ptr(ws["c"][li][t - 1]) if t > 0 else 0, _ptr(ws["c"][li][t]),
                   _ptr(ws["xh"][li][t + 1]), cin + C, cin, 0, 0, 0, ws["Mr"][lv], C, 1.0, eng._s())
        if "gates" in sw:
            rb(Gt)

    def lstm_bwd(li, t, Bn, last):
        ws = eng.ws
        cin, C, lv = LI[li], LS[li], LV[li]
        h, w = eng.H // lv, eng.W // lv
        name = "lstm%d/conv" % (li + 1)
        Gt, dxh = ws["G"][li][t], ws["dxh"][li]
        eng.L.call("pivp_lstm_gates_bwd", _ptr(Gt), _ptr(ws["c"][li][t - 1]) if t > 0 else 0, _ptr(ws["c"][li][t]),
                   _ptr(ws["dln"][li]), 0 if last else _ptr(dxh), cin + C, cin, _ptr(ws["dc"][li]), 0 if last else 1, 0, ws["Mr"][lv], C, eng._s())
        if "dG" in sw:
            rb(Gt)
        dG = View(Gt, 4 * C, 0, 4 * C)
        xsrc = xr[(li, t)] if "xh" in sw else ws["xh"][li][t]
        eng._conv_wgrad(View(xsrc, cin + C, 0, cin + C), Bn, h, w, dG, h, w, 5, 1, 2, eng.g[name + "/W"], eng.g[name + "/b"])
        eng._conv_dgrad(dG, Bn, h, w, eng.p[name + "/W"], None, 5, 1, 2, View(dxh, cin + C, 0, cin + C), h, w)
    eng._lstm_fwd, eng._lstm_bwd = lstm_fwd, lstm_bwd


def run(sw, compute="f32"):
    m = pk.Model(NM, is_cdna=MT == "CDNA", is_dna=MT == "DNA", is_stp=MT == "STP", scheduled_sampling_k=900.0, prefix="t", height=H, width=W, compute=compute)
    p = dict(params)
    if "W" in sw:
        for k_ in p:
            if k_.endswith("/W") and (k_.startswith("lstm") or k_.split("/")[0] in ("enc1", "enc2", "enc4", "enc5", "enc6")):
                p[k_] = torch.from_numpy(np.asarray(p[k_], np.float32)).bfloat16().float().numpy()
    m.load_params(p)
    if compute == "f32" and sw:
        patched(m.engine, sw)
    np.random.seed(99); loss = m([torch.from_numpy(a) for a in batch], 6000); m.cleargrads(); m.backward(); torch.cuda.synchronize()
    g = m.grads
    errs = {}
    for key, v in ref["P"].items():
        r = np.zeros_like(v.data) if v.grad is None else v.grad
        if np.linalg.norm(r) == 0:
            continue
        errs[key] = np.linalg.norm(g[key].astype(np.float64) - r) / np.linalg.norm(r)
    fr = max(np.linalg.norm(m.gen_images[t].double().cpu().numpy() - ref["gen_images"][t].data) / np.linalg.norm(ref["gen_images"][t].data) for t in range(T - 1))
    return float(loss), fr, errs


variants = [("fp32 (no rounding)", set(), "f32"), ("gates", {"gates"}, "f32"), ("dG", {"dG"}, "f32"), ("xh", {"xh"}, "f32"), ("W", {"W"}, "f32"),
            ("all four", {"gates", "dG", "xh", "W"}, "f32"), ("bf16 engine (tcgen05)", set(), "bf16")]
print("%s 64x64 B=%d T=%d, per-tensor gradient relative L2 error vs the float64 oracle" % (MT, B, T))
print("| rounding switched on | frames rel L2 (worst t) | worst tensor | its error | median over tensors | lstm1/conv/W | lstm5/conv/W | enc0/W | masks/W |")
print("|---|---:|---|---:|---:|---:|---:|---:|---:|")
for name, sw, comp in variants:
    loss, fr, errs = run(sw, comp)
    worst = max(errs, key=errs.get)
    print("| %s | %.2e | %s | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f |" % (name, fr, worst, errs[worst], float(np.median(list(errs.values()))),
          errs["lstm1/conv/W"], errs["lstm5/conv/W"], errs["enc0/W"], errs["masks/W"]))
