import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for cfg in ["2 64 64 64 128", "4 32 32 64 128", "4 8 8 192 512", "2 16 16 128 256", "6 16 16 96 256"]:
        r = subprocess.run([sys.executable, __file__] + cfg.split(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
        out = r.stdout.decode().strip().splitlines()
        msg = [l for l in out if "error" in l.lower() or "OK" in l]
        print(cfg, "->", (msg[0] if msg else out[-1])[:160])
    sys.exit(0)
sys.path.insert(0, ROOT)
import numpy as np, torch
import pivp_b200 as pk
SB, H, W, Cx, N4 = map(int, sys.argv[1:6])
L = pk.lib(); s = torch.cuda.current_stream().cuda_stream
rs = np.random.RandomState(3)
P = SB * H * W
Kp = (Cx + 63) // 64 * 64
xh = torch.zeros(P, Kp, device="cuda"); xh[:, :Cx] = torch.from_numpy(rs.standard_normal((P, Cx)).astype(np.float32)).cuda()
xh_b = xh.bfloat16()
dg_b = torch.from_numpy((rs.standard_normal((P, N4)) * 0.1).astype(np.float32)).cuda().bfloat16()
nb = L.query("pivp_tc_wgrad_workspace_bytes", SB, H, W, Cx, N4)
ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
dW = torch.zeros(N4, 25, Cx, device="cuda")
L.call("pivp_tc_wgrad5x5", dg_b.data_ptr(), xh_b.data_ptr(), Kp, SB, H, W, Cx, N4, dW.data_ptr(), ws.data_ptr(), nb, s)
torch.cuda.synchronize()
ref = torch.zeros(N4, 25, Cx, device="cuda")
xf, gf = xh_b[:, :Cx].float().contiguous(), dg_b.float().contiguous()
L.call("pivp_conv2d_wgrad", xf.data_ptr(), Cx, 0, SB, H, W, Cx, gf.data_ptr(), N4, 0, H, W, N4, 5, 5, 1, 2, ref.data_ptr(), 0, s)
torch.cuda.synchronize()
err = float((dW - ref).abs().max() / ref.abs().max())
print("OK relerr %.3e" % err, "center tap err %.3e" % float((dW[:,12]-ref[:,12]).abs().max()/ref.abs().max()))
