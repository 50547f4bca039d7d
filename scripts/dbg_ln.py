import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import pivp_b200 as pk
L = pk.lib()
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
def run(B, HW, C, cs, co, relu, two, scalar):
    os.environ["PIVP_LN_SCALAR"] = "1" if scalar else "0"
    n = HW * C
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    xb = torch.randn(B * HW, cs, device="cuda", generator=g) * 2 + 0.5
    g1 = torch.randn(B * HW, cs, device="cuda", generator=g)
    g2 = torch.randn(B * HW, cs, device="cuda", generator=g)
    ga = 1 + 0.2 * torch.randn(n, device="cuda", generator=g); be = 0.2 * torch.randn(n, device="cuda", generator=g)
    y = torch.zeros(B * HW, C, device="cuda"); dx = torch.zeros(B * HW, C, device="cuda")
    stats = torch.zeros(B, 2, device="cuda")
    dga = torch.zeros(n, device="cuda"); dbe = torch.zeros(n, device="cuda")
    nb = L.query("pivp_layernorm_workspace_bytes", B, n)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device="cuda")
    L.call("pivp_layernorm_fwd", xb.data_ptr(), cs, co, ga.data_ptr(), be.data_ptr(), B, HW, C, 1e-6, y.data_ptr(), C, 0, 0, 0, 0, 0, 0, 0, relu,
           stats.data_ptr(), ws.data_ptr(), ws.numel(), st)
    L.call("pivp_layernorm_bwd", xb.data_ptr(), cs, co, g1.data_ptr(), cs, co, g2.data_ptr() if two else 0, cs, co, ga.data_ptr(), be.data_ptr(),
           stats.data_ptr(), B, HW, C, relu, dx.data_ptr(), C, 0, dga.data_ptr(), dbe.data_ptr(), ws.data_ptr(), ws.numel(), st)
    torch.cuda.synchronize()
    return [t.double().cpu() for t in (y, stats, dx, dga, dbe)]
for (B, HW, C, cs, co, relu, two) in [(2, 4096, 64, 64, 0, 1, 0), (2, 256, 64, 192, 128, 0, 0), (2, 1024, 32, 128, 96, 0, 0), (2, 64, 128, 192, 64, 0, 0), (2, 1024, 32, 32, 0, 1, 1), (3, 256, 32, 64, 32, 0, 0), (3, 256, 32, 32, 0, 1, 1), (3, 64, 64, 96, 32, 0, 0), (3, 16, 128, 192, 64, 0, 0), (3, 1024, 64, 64, 0, 1, 0), (32, 1024, 32, 64, 32, 0, 0)]:
    a = run(B, HW, C, cs, co, relu, two, True); b = run(B, HW, C, cs, co, relu, two, False)
    print((B, HW, C, cs, co, relu, two), ["%.2e" % float((u - v).abs().max() / (u.abs().max() + 1e-30)) for u, v in zip(a, b)])
