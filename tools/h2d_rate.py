"""dev tool (GPU box): pinned host -> device copy rate of one large buffer vs chunked copies (what bounds `e2e`)."""
import time
import torch

n = 1_474_560_000 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for chunks in (1, 4, 16, 64):
    sz = n // chunks
    for _ in range(2):
        for c in range(chunks):
            d[c * sz:(c + 1) * sz].copy_(h[c * sz:(c + 1) * sz], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for c in range(chunks):
            d[c * sz:(c + 1) * sz].copy_(h[c * sz:(c + 1) * sz], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunks {chunks:3d}: {n * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms)")
