"""One eager b32 training step (CDNA 64x64, T=10, bf16) between cudaProfilerStart / cudaProfilerStop, after two warm-up steps:
the program `ncu --profile-from-start off ...` profiles (launch list and --set full captures under profiles/).
    python scripts/ncu_step.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T, H, W = 10, 64, 64
model = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="p", compute="bf16")
opt = pk.Adam().setup(model)
host = [torch.from_numpy(a) for a in pk.concat_examples(pk.data.synthetic_sequences(B, T, H, W))]
step = pk.TrainStep(model, opt, B, T, graph=False)
step.load_batch(*host)
np.random.seed(0)
for i in range(2):
    step(6000 + i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step(6002)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss %.6f, launches per step %d" % (float(loss), step.launches_per_step))
