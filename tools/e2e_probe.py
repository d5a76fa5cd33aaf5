"""Where does the end-to-end step time go?  Host time per submit call, device time per step, for each lag / input kind."""
import sys, time
sys.path.insert(0, ".")
import torch
import tcresnet_b200  # noqa
from tcresnet_b200.engine import Engine, HostFeed

n = 512
eng = Engine(max_batch=n)
params, slots, moving = eng.new_variables(seed=0)
gen = torch.Generator().manual_seed(1)
f32 = [(torch.rand(n, 16000, generator=gen) * 2 - 1).pin_memory() for _ in range(4)]
pcm = [(w * 32767).round().to(torch.int16).pin_memory() for w in f32]
hot = [torch.nn.functional.one_hot(torch.randint(0, 12, (n,), generator=gen), 12).float().pin_memory() for _ in range(4)]
d_w = [w.cuda() for w in f32]
d_h = [h.cuda() for h in hot]
losses = torch.zeros(2, device="cuda")

def device_loop(steps=200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        eng.train_step(d_w[i % 4], d_h[i % 4], params, slots, moving, 0.1, 0.9, 1e-3, dropout_seed=i, losses=losses)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"device loop: host enqueue {1e3*(t1-t0)/steps:.3f} ms/step, total {1e3*(t2-t0)/steps:.3f} ms/step")

def feed_loop(bufs, lag, steps=200, tag=""):
    feed = HostFeed(eng, lag=lag)
    for i in range(4):
        feed.submit(bufs[i % 4], hot[i % 4], params, slots, moving, 0.1, 0.9, 1e-3, dropout_seed=i)
    feed.flush()
    torch.cuda.synchronize()
    ts = []
    t0 = time.perf_counter()
    for i in range(steps):
        a = time.perf_counter()
        feed.submit(bufs[i % 4], hot[i % 4], params, slots, moving, 0.1, 0.9, 1e-3, dropout_seed=i)
        ts.append(time.perf_counter() - a)
    feed.flush()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ts.sort()
    print(f"{tag} lag={lag}: total {1e3*(t2-t0)/steps:.3f} ms/step; submit call median {1e3*ts[len(ts)//2]:.3f} p10 {1e3*ts[len(ts)//10]:.3f} p90 {1e3*ts[9*len(ts)//10]:.3f} ms")

def copy_only(bufs, tag):
    dst = torch.empty_like(bufs[0], device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(50):
        dst.copy_(bufs[i % 4], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    print(f"{tag} copy only: {1e3*dt:.3f} ms/batch = {bufs[0].numel()*bufs[0].element_size()/dt/1e9:.1f} GB/s")

for _ in range(2):
    device_loop()
copy_only(f32, "f32")
copy_only(pcm, "pcm16")
for lag in (0, 1, 2):
    feed_loop(f32, lag, tag="f32  ")
for lag in (0, 1, 2):
    feed_loop(pcm, lag, tag="pcm16")

# pure host cost of one submit: the pipeline is empty, so nothing blocks on the GPU
feed = HostFeed(eng, lag=2)
costs = []
for rep in range(30):
    feed.flush(); torch.cuda.synchronize()
    a = time.perf_counter()
    feed.submit(pcm[0], hot[0], params, slots, moving, 0.1, 0.9, 1e-3, dropout_seed=rep)
    b = time.perf_counter()
    feed.submit(pcm[1], hot[1], params, slots, moving, 0.1, 0.9, 1e-3, dropout_seed=rep)
    c = time.perf_counter()
    costs.append((b - a, c - b))
feed.flush()
costs.sort()
print(f"non-blocking submit (host only): first median {1e3*sorted(x[0] for x in costs)[15]:.3f} ms, second median {1e3*sorted(x[1] for x in costs)[15]:.3f} ms")
t = []
for rep in range(30):
    torch.cuda.synchronize()
    a = time.perf_counter()
    eng.train_step(d_w[0], d_h[0], params, slots, moving, 0.1, 0.9, 1e-3, dropout_seed=rep, losses=losses)
    t.append(time.perf_counter() - a)
t.sort()
print(f"non-blocking device train_step (host only): median {1e3*t[15]:.3f} ms  min {1e3*t[0]:.3f}")
import os
print("cpus", os.cpu_count(), "loadavg", os.getloadavg())
