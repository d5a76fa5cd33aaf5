"""Debug helper: per-CTA phase timeline of conv_fwd(block2/conv2_0) and dw_grouped (TCR_DEBUG_TIMELINE=1)."""
import os, sys, ctypes as C
import numpy as np, torch
os.environ["TCR_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcresnet_b200
from tcresnet_b200.engine import Engine
from tcresnet_b200 import _lib as L
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="TCResNet8"); ap.add_argument("--width", type=float, default=1.0); ap.add_argument("--batch", type=int, default=512)
A = ap.parse_args()
N = A.batch
eng = Engine(model=A.model, width_multiplier=A.width, max_batch=N)
dev = eng.device
p, s, m = eng.new_variables(0)
wav = torch.rand(N, 16000, device=dev) * 2 - 1
hot = torch.nn.functional.one_hot(torch.randint(0, 12, (N,), device=dev), 12).float()
for i in range(5):
    eng.train_step(wav, hot, p, s, m, 0.1)
torch.cuda.synchronize()
ptr, numel = C.c_void_p(), C.c_int64()
L.check(eng.lib, eng.lib.tcr_workspace_tensor(eng._h, b"timeline", C.byref(ptr), C.byref(numel)), "timeline")
buf = torch.empty(8 * 8192, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
rt = C.CDLL('libcudart.so.12')
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
assert rt.cudaMemcpy(buf.data_ptr(), ptr.value, 8 * 8192 * 8, 3) == 0
t = buf.cpu().numpy().reshape(8192, 8)
f = t[:1024, :7].astype(np.float64)
f = f[f[:, 0] > 0]
t0 = f[:, 0].min()
print("conv_fwd block2/conv2_0: CTAs", len(f))
names = ["start", "tile staged", "weights landed+sync", "compute done", "y stored", "stats done", "end"]
for i, nme in enumerate(names):
    col = f[:, i] - t0
    print(f"  {nme:22s} min {col.min()/1e3:7.2f} us  median {np.median(col)/1e3:7.2f}  max {col.max()/1e3:7.2f}")
p = t[2048:2048 + 1024, :8].astype(np.float64)
p = p[p[:, 0] > 0]
g = t[:1024, :8].astype(np.float64); g = g[g[:, 0] > 0]
if len(p) and len(g):
    base = p[:, 0].min()
    print("producer block1/conv1_1 -> consumer block2/conv2_0 (us from the producer's first CTA start):")
    print(f"  producer: stats done max {(p[:,5].max()-base)/1e3:.2f}  end (after cluster publish) median {(np.median(p[:,6])-base)/1e3:.2f} max {(p[:,6].max()-base)/1e3:.2f}")
    print(f"  consumer: CTA start min {(g[:,0].min()-base)/1e3:.2f} median {(np.median(g[:,0])-base)/1e3:.2f} max {(g[:,0].max()-base)/1e3:.2f};  dependency released min {(g[:,7].min()-base)/1e3:.2f} median {(np.median(g[:,7])-base)/1e3:.2f} max {(g[:,7].max()-base)/1e3:.2f};  tile staged median {(np.median(g[:,1])-base)/1e3:.2f}")
d = t[4096:4096 + 3000]
d = d[d[:, 0] > 0]
t0 = d[:, 0].min()
print("dw_grouped: CTAs", len(d), " span", (d[:, 2].max() - t0) / 1e3, "us")
dur = (d[:, 2] - d[:, 0]) / 1e3
start = (d[:, 0] - t0) / 1e3
for l in sorted(set(d[:, 3])):
    sel = d[:, 3] == l
    print(f"  layer {int(l):2d}: ctas {sel.sum():3d}  start {start[sel].min():6.1f}..{start[sel].max():6.1f} us  duration min {dur[sel].min():6.1f} med {np.median(dur[sel]):6.1f} max {dur[sel].max():6.1f}")
