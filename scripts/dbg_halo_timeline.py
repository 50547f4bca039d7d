"""Per-CTA timeline of the halo-patch tcgen05 convolution (clock64 stamps), for one ConvLSTM-shaped launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk
L = pk.lib()
st = torch.cuda.current_stream().cuda_stream
def run(B, H, W, Kc, C, mode, N=None, BN=128):
    M = B * H * W
    N = 4 * C if mode == 1 else N
    x = torch.randn(M, Kc, device="cuda").bfloat16()
    w = (torch.randn(N, 25, Kc, device="cuda") / (25 * Kc) ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    gates = torch.empty(M, N, device="cuda"); cp = torch.randn(M, C, device="cuda"); co = torch.empty(M, C, device="cuda")
    h = torch.empty(M, Kc, device="cuda"); hb = torch.empty(M, Kc, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda")
    dbg = torch.zeros(4096, 8, dtype=torch.int64, device="cuda")
    def call():
        if mode == 1:
            L.call("pivp_tc_conv5x5", x.data_ptr(), Kc, B, H, W, Kc, w.data_ptr(), N, 128, 1, bias.data_ptr(), 0, 0, 0,
                   gates.data_ptr(), cp.data_ptr(), co.data_ptr(), h.data_ptr(), Kc, Kc - C, hb.data_ptr(), Kc, Kc - C, 0, 0, 0, C, 1.0, 0, 0, st)
        else:
            L.call("pivp_tc_conv5x5", x.data_ptr(), Kc, B, H, W, Kc, w.data_ptr(), N, BN, 0, 0, out.data_ptr(), N, 0,
                   0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, C, 0.0, 0, 0, st)
    for _ in range(3): call()
    torch.cuda.synchronize()
    L.call("pivp_tc_set_debug_buffer", dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    L.call("pivp_tc_set_debug_buffer", 0)
    d = dbg.cpu().numpy(); d = d[d[:, 0] > 0]
    rel = (d[:, 1:6] - d[:, :1]).astype(np.float64)
    names = ["setup", "first data", "last MMA issued", "accum ready", "epilogue done"]
    print("B%d %dx%d Kc=%d N=%d mode%d: %.1f us, %d CTAs; cycles since CTA start (mean / max):" % (B, H, W, Kc, N, mode, e0.elapsed_time(e1) * 1e3, len(d)))
    for i, n in enumerate(names):
        print("   %-16s %8.0f %8.0f" % (n, rel[:, i].mean(), rel[:, i].max()))
    print("   CTA start spread %8.0f cycles" % (d[:, 0].max() - d[:, 0].min()))
run(32, 32, 32, 64, 32, 1)
run(32, 32, 32, 128, 32, 1)
run(32, 16, 16, 128, 64, 1)
run(32, 32, 32, 128, 32, 0, N=64, BN=64)
run(32, 16, 16, 256, 64, 0, N=192, BN=192)
run(32, 16, 16, 256, 64, 0, N=192, BN=96)      # lstm6 input gradient as the engine launches it (two N tiles)
run(32, 8, 8, 192, 128, 1)                     # lstm5 forward: two-image pair geometry
run(32, 8, 8, 512, 128, 0, N=192, BN=32)       # lstm5 input gradient
