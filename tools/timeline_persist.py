"""Debug helper: per-phase timeline of the persistent step kernel (TCR_DEBUG_TIMELINE=1)."""
import os, sys, ctypes as C
import numpy as np, torch
os.environ["TCR_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcresnet_b200
from tcresnet_b200.engine import Engine
from tcresnet_b200 import _lib as L
eng = Engine(max_batch=512)
dev = eng.device
p, s, m = eng.new_variables(0)
wav = torch.rand(512, 16000, device=dev) * 2 - 1
hot = torch.nn.functional.one_hot(torch.randint(0, 12, (512,), device=dev), 12).float()
for i in range(5):
    eng.train_step(wav, hot, p, s, m, 0.1)
torch.cuda.synchronize()
ptr, numel = C.c_void_p(), C.c_int64()
L.check(eng.lib, eng.lib.tcr_workspace_tensor(eng._h, b"timeline", C.byref(ptr), C.byref(numel)), "timeline")
N = 16 * 8192
buf = torch.empty(N, dtype=torch.int64, device=dev)
rt = C.CDLL('libcudart.so.12'); rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
assert rt.cudaMemcpy(buf.data_ptr(), ptr.value, N * 8, 3) == 0
t = buf.cpu().numpy()[:64 * 512 * 2].reshape(64, 512, 2).astype(np.float64)
t0 = t[0, :, 0][t[0, :, 0] > 0].min()
kinds = ["transpose", "fwd", "fin_fwd", "head", "bwd", "fin_bwd", "dw", "grad"]
print("ph  start(first)  work(min/med/max us)  all-done  next-phase-start  barrier")
prev_done = None
for ph in range(64):
    a = t[ph]
    sel = a[:, 0] > 0
    if not sel.any(): break
    st, en = a[sel, 0] - t0, a[sel, 1] - t0
    w = (en - st) / 1e3
    nxt = t[ph + 1][:, 0]; nxt = nxt[nxt > 0]
    nstart = (nxt.min() - t0) / 1e3 if len(nxt) else float('nan')
    print(f"{ph:2d}  {st.min()/1e3:8.1f}   work {w.min():6.1f} {np.median(w):6.1f} {w.max():6.1f}   done {en.max()/1e3:8.1f}   next {nstart:8.1f}   barrier {nstart - en.max()/1e3:5.1f}")

f = buf.cpu().numpy()[12 * 8192: 12 * 8192 + 256 * 8].reshape(256, 8)[:, :7].astype(np.float64)
f = f[f[:, 0] > 0]
t0 = f[:, 0].min()
print("in-body stamps, fwd block2/conv2_0 inside the persistent kernel: CTAs", len(f))
for i, nme in enumerate(["start", "tile staged", "weights landed+sync", "compute done", "y stored", "stats done", "end"]):
    col = f[:, i] - t0
    print(f"  {nme:22s} min {col.min()/1e3:7.2f} us  median {np.median(col)/1e3:7.2f}  max {col.max()/1e3:7.2f}")
