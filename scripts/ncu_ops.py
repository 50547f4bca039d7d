"""The fused transform ops alone (CDNA / DNA forward and backward at one batch size), for ncu.
    python scripts/ncu_ops.py [batch=256]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pivp_b200 as pk
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = W = 64
L = pk.lib()
dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
prev, g = torch.rand(B, 3, H, W, device=dev), torch.randn(B, 3, H, W, device=dev)
# CDNA
e, a, k = torch.randn(B, 3, H, W, device=dev), 2 * torch.randn(B, 11, H, W, device=dev), torch.randn(B, 250, device=dev)
out, de, da, dk = torch.empty_like(prev), torch.empty_like(e), torch.empty_like(a), torch.empty_like(k)
nb = L.query("pivp_cdna_fused_bwd_workspace_bytes", B, H, W, 10)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
# DNA
e2, a2 = torch.randn(B, 25, H, W, device=dev), 2 * torch.randn(B, 2, H, W, device=dev)
de2, da2 = torch.empty_like(e2), torch.empty_like(a2)
nb2 = L.query("pivp_dna_fused_bwd_workspace_bytes", B, H, W)
ws2 = torch.empty(max(nb2, 16), dtype=torch.uint8, device=dev)
for rep in range(3):
    if rep == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    L.call("pivp_cdna_fused_fwd", prev.data_ptr(), e.data_ptr(), a.data_ptr(), k.data_ptr(), out.data_ptr(), B, H, W, 10, st())
    L.call("pivp_cdna_fused_bwd", g.data_ptr(), prev.data_ptr(), e.data_ptr(), a.data_ptr(), k.data_ptr(), de.data_ptr(), da.data_ptr(), dk.data_ptr(), 0, 0,
           B, H, W, 10, ws.data_ptr(), nb, st())
    L.call("pivp_dna_fused_fwd", prev.data_ptr(), e2.data_ptr(), a2.data_ptr(), out.data_ptr(), B, H, W, st())
    L.call("pivp_dna_fused_bwd", g.data_ptr(), prev.data_ptr(), e2.data_ptr(), a2.data_ptr(), de2.data_ptr(), da2.data_ptr(), 0, 0, B, H, W, ws2.data_ptr(), nb2, st())
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok")
