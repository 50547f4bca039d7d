"""Generate the frozen oracle vectors under tests/golden/ (run from the repo root).

The reference cannot run here (Chainer 2.0.1 / Python 2 unavailable), so these
vectors are produced by the oracle itself AFTER it was cross-checked against the torch
restatement; they freeze its behaviour so later edits to ``oracle/`` are caught, and they
travel to the GPU box where the CUDA path is compared with them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import model as M  # noqa: E402

CASES = {
    # name: (model_type, num_masks, schedsamp_k, iter_num, B, T, H)
    "cdna_sched": ("CDNA", 10, 900.0, 6000, 3, 5, 64),
    "cdna_feedself": ("CDNA", 10, -1.0, 0, 2, 4, 64),
    "dna_sched": ("DNA", 1, 900.0, 6000, 3, 4, 64),
    "stp_sched": ("STP", 10, 900.0, 6000, 3, 4, 64),
}


def run_case(name):
    mt, nm, k, it, B, T, H = CASES[name]
    cfg = M.Config(mt, nm, schedsamp_k=k, height=H, width=H)
    params = M.init_params(cfg, seed=4321)
    rs = np.random.RandomState(7)
    for key in sorted(params):            # move LN/bias params off their 1/0 init so their grads matter
        if not key.endswith("/W"):
            params[key] = (params[key] + 0.05 * rs.standard_normal(params[key].shape)).astype(np.float32)
    batch = M.concat_examples(M.synthetic_sequences(B, T, cfg, seed=1234))
    np.random.seed(99)
    take = []
    out = M.forward(params, batch, it, cfg, take_gt_log=take)
    M.G.backward(out["loss"])
    res = {
        "loss": np.float32(out["loss"].data),
        "psnr_all": np.float32(out["psnr_all"]),
        "recon_costs": np.array(out["recon_costs"], np.float32),
        "n_gt": np.int32(-1 if out["n_gt"] is None else out["n_gt"]),
        "take_gt": np.array(take, dtype=np.bool_).reshape(len(take), B),
        "gen_last": out["gen_images"][-1].data.astype(np.float32),
        "gen_first": out["gen_images"][0].data[:1].astype(np.float32),
        "masks_last": out["trace"][-1]["masks"].data[:1].astype(np.float16),
        "gen_state_last": out["gen_states"][-1].data.astype(np.float32),
    }
    for key, v in out["P"].items():
        g = np.zeros_like(v.data) if v.grad is None else v.grad
        res["gsum/" + key] = np.float64(g.astype(np.float64).sum())
        res["gl2/" + key] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
    return res


if __name__ == "__main__":
    for name in CASES:
        res = run_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, float(res["loss"]), int(res["n_gt"]))
