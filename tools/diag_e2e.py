"""Diagnostic: where does the host-buffer path spend its time? (run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_key_estimation_b200 as ake
from audio_key_estimation_b200 import _lib, synth

B, n = int(os.environ.get("B", 256)), 48000 * 30
w = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_seed0.npz"))
net = ake.PitchClassNet(288, 12, 2, 7, opt=ake.default_opt(genre=True))
net.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
est = ake.KeyEstimator(net.cuda().eval(), 48000)
host = torch.empty((B, n), dtype=torch.float32).pin_memory()
host[:8] = synth.synth_batch(0, 8, n, 48000)
for i in range(8, B):
    host[i] = host[i % 8]
dev = torch.empty((B, n), dtype=torch.float32, device="cuda")
for _ in range(3):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D {host.numel()*4/1e9:.2f} GB: {dt*1e3:.2f} ms -> {host.numel()*4/dt/1e9:.1f} GB/s")
out = None
for _ in range(2):
    out = est.estimate_host(host, out=out)
_lib.profile_enable(True)
t0 = time.perf_counter()
for _ in range(5):
    out = est.estimate_host(host, out=out)
dt = (time.perf_counter() - t0) / 5
prof = _lib.profile_collect()
_lib.profile_enable(False)
print(f"estimate_host: {dt*1e3:.2f} ms/step -> {B/dt:.0f} clips/s; sections per step: " +
      ", ".join(f"{k}={v[0]/5:.2f}ms" for k, v in sorted(prof.items())))
for _ in range(2):
    est.estimate_device(dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    est.estimate_device(dev)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"estimate_device: {dt*1e3:.2f} ms/step")
