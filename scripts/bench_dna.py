import sys, os, json
sys.path.insert(0, '/root/repo')
import torch, pivp_b200 as pk
L = pk.lib(); dev="cuda"; H=W=64
def op_time(fn, nsets, iters=20):
    for i in range(nsets): fn(i, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters): fn(i % nsets, torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3 / iters)
    return best
HW=H*W
byt = {"fwd": 33*HW*4, "bwd": 60*HW*4}
for Bc in (32, 256):
    nsets = max(2, int(600e6 // (Bc * 60 * HW * 4)) + 1)
    sets = [dict(prev=torch.rand(Bc,3,H,W,device=dev), e=torch.randn(Bc,25,H,W,device=dev), a=2*torch.randn(Bc,2,H,W,device=dev), g=torch.randn(Bc,3,H,W,device=dev)) for _ in range(nsets)]
    for t in sets: t.update(out=torch.empty_like(t["prev"]), de=torch.empty_like(t["e"]), da=torch.empty_like(t["a"]))
    nb = L.query("pivp_dna_fused_bwd_workspace_bytes", Bc, H, W); wsb = torch.empty(max(nb,16), dtype=torch.uint8, device=dev)
    def fwd(i, st): t=sets[i]; L.call("pivp_dna_fused_fwd", t["prev"].data_ptr(), t["e"].data_ptr(), t["a"].data_ptr(), t["out"].data_ptr(), Bc, H, W, st)
    def bwd(i, st): t=sets[i]; L.call("pivp_dna_fused_bwd", t["g"].data_ptr(), t["prev"].data_ptr(), t["e"].data_ptr(), t["a"].data_ptr(), t["de"].data_ptr(), t["da"].data_ptr(), 0, 0, Bc, H, W, wsb.data_ptr(), nb, st)
    for nm, fn in (("fwd", fwd), ("bwd", bwd)):
        tm = op_time(fn, nsets)
        print("DNA b%d %s: %.2f us, %.0f GB/s, frac %.3f" % (Bc, nm, tm*1e6, byt[nm]*Bc/tm/1e9, byt[nm]*Bc/tm/1e9/6554.2))
