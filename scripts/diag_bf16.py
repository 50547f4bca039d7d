"""Diagnostic: per-tensor error of the bf16 engine vs the float64 oracle and vs the fp32 engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk
from oracle import model as OM, npgrad as G

mt, nm, k = sys.argv[1] if len(sys.argv) > 1 else "CDNA", 10, 900.0
if mt == "DNA": nm = 1
H = W = 64; B, T = 2, 4
cfg = OM.Config(mt, nm, schedsamp_k=k, height=H, width=W, dtype=np.float64)
params = OM.init_params(cfg)
if os.environ.get("PERTURB"):                      # the parameter set of tests/test_gpu_tc.py::test_model_bf16_within_tolerance_of_oracle
    rs = np.random.RandomState(7)
    for key in sorted(params):
        if not key.endswith("/W"):
            params[key] = params[key] + 0.05 * rs.standard_normal(params[key].shape)
batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
np.random.seed(99); ref = OM.forward(params, batch, 6000, cfg); G.backward(ref["loss"])
res = {}
for mode in ("f32", "bf16"):
    m = pk.Model(nm, is_cdna=(mt == "CDNA"), is_dna=(mt == "DNA"), is_stp=(mt == "STP"), scheduled_sampling_k=k, prefix="t", height=H, width=W, compute=mode)
    m.load_params(params)
    np.random.seed(99); loss = m([torch.from_numpy(a) for a in batch], 6000); m.cleargrads(); m.backward(); torch.cuda.synchronize()
    res[mode] = (float(loss), [g.cpu().numpy() for g in m.gen_images], m.grads)
    print(mode, "loss", float(loss), "oracle", float(ref["loss"].data))
    for t in range(T - 1):
        a, b = res[mode][1][t].astype(np.float64), ref["gen_images"][t].data
        print("  gen[%d] max-rel %.3e  l2-rel %.3e" % (t, np.abs(a - b).max() / np.abs(b).max(), np.linalg.norm(a - b) / np.linalg.norm(b)))
print("%-28s %10s %10s %10s | %10s" % ("tensor", "maxrel", "l2rel", "cos", "f32 maxrel"))
for key, v in sorted(ref["P"].items()):
    r = np.zeros_like(v.data) if v.grad is None else v.grad
    g = res["bf16"][2][key].astype(np.float64); g32 = res["f32"][2][key].astype(np.float64)
    den = np.abs(r).max() + 1e-30
    cos = (g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
    print("%-28s %10.3e %10.3e %10.6f | %10.3e" % (key, np.abs(g - r).max() / den, np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30), cos, np.abs(g32 - r).max() / den))
