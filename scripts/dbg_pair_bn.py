"""ConvLSTM-5 (8x8 maps) input gradient: N-tile sweep of the halo pair geometry vs the per-tap kernel, each as a 20-launch CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pivp_b200 as pk
L = pk.lib()
B, H, W, Kc, N = 32, 8, 8, 512, 192
M = B * H * W
x = torch.randn(M, Kc, device="cuda").bfloat16()
w = (torch.randn(N, 25, Kc, device="cuda") / (25 * Kc) ** 0.5).bfloat16()
out = torch.empty(M, N, device="cuda")
def t(BN):
    def call(st):
        L.call("pivp_tc_conv5x5", x.data_ptr(), Kc, B, H, W, Kc, w.data_ptr(), N, BN, 0, 0, out.data_ptr(), N, 0,
               0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 128, 0.0, 0, 0, st)
    call(torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): call(torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / 20)
    return best
print("PIVP_TC_HALO=%s SPLITK=%s" % (os.environ.get("PIVP_TC_HALO", "1"), os.environ.get("PIVP_TC_HALO_SPLITK", "auto")), {bn: round(t(bn), 1) for bn in (32, 48, 64, 96, 192)})
