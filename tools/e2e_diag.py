"""Diagnostic: per-kernel event times inside the host-buffer step loop (fp32 / pcm16 / pcm16 + device input stage)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcresnet_b200  # noqa
from tcresnet_b200.engine import Engine, HostFeed
from tcresnet_b200.datasets import device_input_stage as D
n = 512
eng = Engine(max_batch=n, dropout_keep_prob=0.5)
dev = eng.device
p, s, m = eng.new_variables(0)
gen = torch.Generator(device=dev).manual_seed(1)
wavs = [torch.rand(n, 16000, device=dev, generator=gen) * 2 - 1 for _ in range(4)]
hots = [torch.nn.functional.one_hot(torch.randint(0, 12, (n,), device=dev), 12).float().cpu().pin_memory() for _ in range(4)]
h_wavs = [w.cpu().pin_memory() for w in wavs]
h_pcm = [(w.clamp(-1, 1) * 32767.0).round().to(torch.int16).cpu().pin_memory() for w in wavs]
rs = np.random.RandomState(99)
stage = D.DeviceInputStage(eng, [rs.uniform(-0.5, 0.5, 960000).astype(np.float32) for _ in range(6)])
h_clips = [torch.from_numpy(np.frombuffer(D.draw_clips(rs, [16000] * n, rs.uniform(size=n) < 0.1, 16000, stage.bg_lengths).tobytes(), np.uint8).copy()).pin_memory() for _ in h_pcm]
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
for rep in range(3):
    for name, bufs, clips in (("fp32", h_wavs, None), ("pcm16", h_pcm, None), ("aug", h_pcm, h_clips)):
        feed = HostFeed(eng, lag=2)
        for i in range(20):
            feed.submit(bufs[i % 4], hots[i % 4], p, s, m, 0.1, dropout_seed=i, h_clips=clips[i % 4] if clips else None,
                        background=stage.background if clips else None)
        feed.flush()
        torch.cuda.synchronize()
        eng.profile(True)
        t0 = time.perf_counter()
        for i in range(100):
            feed.submit(bufs[i % 4], hots[i % 4], p, s, m, 0.1, dropout_seed=i, h_clips=clips[i % 4] if clips else None,
                        background=stage.background if clips else None)
        feed.flush()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = eng.profile_read()
        eng.profile(False)
        print(rep, name, f"{n*100/dt/1e3:.0f} k utt/s  {dt*10:.3f} ms/step  ", " ".join(f"{k}:{v[0]/v[1]*1e3:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1][0])[:7]))
