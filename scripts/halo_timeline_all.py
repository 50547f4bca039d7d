"""clock64 timeline (per-CTA stamps) of the 14 ConvLSTM tcgen05 launches of the b32 step, with the production arguments
(bf16 gate storage, fused LayerNorm partials, the engine's N tiles / split-K).  Prints a markdown table.
    python scripts/halo_timeline_all.py [--batch 32]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=32); args = ap.parse_args()
L = pk.lib()
B = args.batch
LS, LI, LV = (32, 32, 64, 64, 128, 64, 32), (32, 32, 32, 64, 64, 128, 96), (2, 2, 4, 4, 8, 4, 2)
print("cycles since CTA start: mean (max) over CTAs; graph columns: 8 back-to-back launches of the same call in one CUDA graph, %globaltimer per CTA")
print("| launch | grid | setup | first data | last MMA | accum ready | epilogue done | MMA floor (cyc) | graph: period us | kernel span us | gap us | start spread us | end spread us |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
def run(name, H, W, Kc, C, mode, N, BN, ln):
    M = B * H * W
    x = torch.randn(M, Kc, device="cuda").bfloat16()
    w = (torch.randn(N, 25, Kc, device="cuda") / (25 * Kc) ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    gates = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); cp = torch.randn(M, C, device="cuda"); co = torch.empty(M, C, device="cuda")
    h = torch.empty(M, Kc, device="cuda"); hb = torch.empty(M, Kc, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda")
    part = torch.zeros(B * 64 * 2 + 16, device="cuda")
    n = H * W * C
    gamma, beta = torch.ones(n, device="cuda"), torch.zeros(n, device="cuda")
    y, yb = torch.empty(M, C, device="cuda"), torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
    stats, counter = torch.zeros(B, 2, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
    dbg = torch.zeros(4096, 8, dtype=torch.int64, device="cuda")
    def call():
        st = torch.cuda.current_stream().cuda_stream          # looked up per call: under graph capture the current stream is the capture stream
        if mode == 1 and ln and os.environ.get("PIVP_TC_FUSE_LN", "0") != "0":        # the production launch: cell + LayerNorm in one kernel
            L.call("pivp_tc_conv5x5_ln", x.data_ptr(), Kc, B, H, W, Kc, w.data_ptr(), C, bias.data_ptr(), gates.data_ptr(), cp.data_ptr(), co.data_ptr(),
                   h.data_ptr(), Kc, Kc - C, hb.data_ptr(), Kc, Kc - C, 1.0, 0, part.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6,
                   y.data_ptr(), C, 0, yb.data_ptr(), 64, 0, stats.data_ptr(), counter.data_ptr(), st)
        elif mode == 1:
            L.call("pivp_tc_conv5x5", x.data_ptr(), Kc, B, H, W, Kc, w.data_ptr(), N, 128, 1, bias.data_ptr(), 0, 0, 0,
                   gates.data_ptr(), cp.data_ptr(), co.data_ptr(), h.data_ptr(), Kc, Kc - C, hb.data_ptr(), Kc, Kc - C, 0, 0, 0, C, 1.0, 2,
                   part.data_ptr() if ln else 0, st)
        else:
            L.call("pivp_tc_conv5x5", x.data_ptr(), Kc, B, H, W, Kc, w.data_ptr(), N, BN, 0, 0, out.data_ptr(), N, 0,
                   0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, C, 0.0, 0, 0, st)
    for _ in range(3): call()
    torch.cuda.synchronize()
    dbg = torch.zeros(512 * 256, 8, dtype=torch.int64, device="cuda")
    L.call("pivp_tc_set_debug_buffer", dbg.data_ptr())
    call(); torch.cuda.synchronize()
    d = dbg[:256].cpu().numpy(); d = d[d[:, 0] > 0]
    rel = (d[:, 1:6] - d[:, :1]).astype(np.float64)
    # graph of NL dependent launches, each with its own stamp rows
    NL = 8
    dbg.zero_()
    L.call("pivp_tc_set_debug_buffer", dbg.data_ptr())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(NL): call()
    L.call("pivp_tc_set_debug_buffer", 0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    period = e0.elapsed_time(e1) * 1e3 / NL
    dd = dbg[:NL * 256].cpu().numpy().reshape(NL, 256, 8)
    st0, st1, en0, en1 = [], [], [], []
    for k in range(NL):
        r = dd[k][dd[k][:, 6] > 0]
        st0.append(r[:, 6].min()); st1.append(r[:, 6].max()); en0.append(r[:, 7].min()); en1.append(r[:, 7].max())
    st0, st1, en0, en1 = [np.array(v, np.float64) for v in (st0, st1, en0, en1)]
    span = (en1 - st0)[1:].mean() * 1e-3
    gap = (st0[1:] - en1[:-1]).mean() * 1e-3
    floor = 25 * (Kc // 64) * 4 * (BN if mode == 0 else 128) / 2
    f = lambda i: "%.0f (%.0f)" % (rel[:, i].mean(), rel[:, i].max())
    print("| %s | %d CTAs | %s | %s | %s | %s | %s | %.0f x MS | %.2f | %.2f | %.2f | %.2f | %.2f |" % (name, len(d), f(0), f(1), f(2), f(3), f(4), floor,
          period, span, gap, (st1 - st0)[1:].mean() * 1e-3, (en1 - en0)[1:].mean() * 1e-3))
for li, (c, cin, lv) in enumerate(zip(LS, LI, LV)):
    H = W = 64 // lv
    Kp = (cin + c + 63) // 64 * 64
    cx = cin + c
    run("lstm%d fwd" % (li + 1), H, W, Kp, c, 1, 4 * c, 128, H % 16 == 0)
    bn = 96 if (H == 8 and cx % 96 == 0) else cx
    run("lstm%d dgrad" % (li + 1), H, W, 4 * c, c, 0, cx, bn, False)
