"""Margins of tests/test_gpu_tc.py::test_model_bf16_within_tolerance_of_oracle (bf16 mode, 64x64, B=2, T=4, against the float64 oracle): prints the
worst frame / mask-logit relative L2 and the worst per-tensor gradient relative L2 / cosine of every parametrisation.  python scripts/bf16_margins.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); 

import numpy as np, torch
import __graft_entry__; __graft_entry__.build()
import pivp_b200 as pk
from oracle import model as OM, npgrad as G
def l2rel(a, b):
    a = a.detach().float().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))
for mt, nm, k in [("CDNA", 10, 900.0), ("CDNA", 10, -1.0), ("DNA", 1, 900.0), ("STP", 10, 900.0), ("CDNA", 4, 900.0)]:
    H = W = 64; B, T = 2, 4
    cfg = OM.Config(mt, nm, schedsamp_k=k, height=H, width=W, dtype=np.float64)
    params = OM.init_params(cfg)
    rs = np.random.RandomState(7)
    for key in sorted(params):
        if not key.endswith("/W"):
            params[key] = params[key] + 0.05 * rs.standard_normal(params[key].shape)
    batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
    np.random.seed(99)
    ref = OM.forward(params, batch, 6000, cfg)
    G.backward(ref["loss"])
    model = pk.Model(nm, is_cdna=(mt == "CDNA"), is_dna=(mt == "DNA"), is_stp=(mt == "STP"), scheduled_sampling_k=k, prefix="t", height=H, width=W, compute="bf16")
    model.load_params(params)
    np.random.seed(99)
    loss = model([torch.from_numpy(a) for a in batch], 6000)
    model.cleargrads(); model.backward(); torch.cuda.synchronize()
    fr = max(l2rel(model.gen_images[t], ref["gen_images"][t].data) for t in range(T - 1))
    mk = max(l2rel(model.engine.ws["mask_pre"][t], ref["trace"][t]["mask_pre"].data) for t in range(T - 1))
    worst_e, worst_c = 0, 1
    grads = model.grads
    for key, v in ref["P"].items():
        r = np.zeros_like(v.data) if v.grad is None else v.grad
        g = grads[key].astype(np.float64)
        e = np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
        cos = (g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
        worst_e, worst_c = max(worst_e, e), min(worst_c, cos)
    print(mt, nm, k, "loss rel %.2e frames %.4f masks %.4f grad worst relL2 %.4f cos %.5f" % (abs(float(loss) - float(ref["loss"].data)) / abs(float(ref["loss"].data)), fr, mk, worst_e, worst_c), flush=True)
