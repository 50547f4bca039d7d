"""CDNA fused op micro-benchmark: CUDA-event time per launch, cold L2 (rotating input sets > 126 MB) and warm L2.

usage: python scripts/bench_cdna.py [B ...]    -> one line per batch size, fwd and bwd, GB/s against MEASURED_PEAKS.json
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pivp_b200 as pk

H = W = 64
M = 10
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
L = pk.lib()


def cur():
    return torch.cuda.current_stream().cuda_stream


def run(B, nsets, iters=20):
    fwd_bytes = B * ((3 + 3 + 11 + 3) * H * W * 4 + 1000)
    bwd_bytes = B * ((3 + 3 + 3 + 11 + 3 + 11) * H * W * 4 + 2000)          # no d_prev (training with ground-truth / sampled frames)
    sets = []
    for _ in range(nsets):
        prev = torch.rand(B, 3, H, W, device="cuda"); e = torch.randn(B, 3, H, W, device="cuda")
        a = 2 * torch.randn(B, M + 1, H, W, device="cuda"); k = torch.randn(B, 25 * M, device="cuda"); g = torch.randn(B, 3, H, W, device="cuda")
        out = torch.empty_like(prev); de = torch.empty_like(e); da = torch.empty_like(a); dk = torch.empty_like(k)
        sets.append((prev, e, a, k, g, out, de, da, dk))
    nb = L.query("pivp_cdna_fused_bwd_workspace_bytes", B, H, W, M)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")

    def fwd(s):
        prev, e, a, k, g, out, de, da, dk = s
        L.call("pivp_cdna_fused_fwd", prev.data_ptr(), e.data_ptr(), a.data_ptr(), k.data_ptr(), out.data_ptr(), B, H, W, M, cur())

    def bwd(s):
        prev, e, a, k, g, out, de, da, dk = s
        L.call("pivp_cdna_fused_bwd", g.data_ptr(), prev.data_ptr(), e.data_ptr(), a.data_ptr(), k.data_ptr(), de.data_ptr(), da.data_ptr(),
               dk.data_ptr(), 0, 0, B, H, W, M, ws.data_ptr(), nb, cur())

    res = {}
    for name, fn, nbytes in (("fwd", fwd, fwd_bytes), ("bwd", bwd, bwd_bytes)):
        for i in range(3 * nsets):
            fn(sets[i % nsets])
        torch.cuda.synchronize()
        # graph of `iters` launches walking the sets: launch gaps excluded, cold L2 when nsets * bytes >> 126 MB
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(iters):
                fn(sets[i % nsets])
        gr.replay(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            ev0.record(); gr.replay(); ev1.record(); torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1) * 1e3 / iters)
        res[name] = dict(us=round(best, 2), GBs=round(nbytes / best / 1e3, 1), frac=round(nbytes / best / 1e3 / peak, 3), MB=round(nbytes / 1e6, 2))
    return res


if __name__ == "__main__":
    for B in [int(v) for v in sys.argv[1:]] or [32, 256]:
        per = B * 37 * H * W * 4
        cold = max(1, int(600e6 // per) + 1)
        print(json.dumps({"B": B, "cold_sets": cold, "cold": run(B, cold), "warm": run(B, 1)}))
