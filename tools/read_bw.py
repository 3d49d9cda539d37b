"""Read-only HBM bandwidth ceiling on this box: torch.sum over a 6.3 GB fp32 tensor (cub reduce)."""
import torch
x = torch.ones(1024 * 480 * 640 * 5, device="cuda")
for _ in range(3): x.sum()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): x.sum()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"torch.sum read-only: {x.numel()*4/ms/1e6:.1f} GB/s ({ms:.3f} ms)")
y = torch.empty_like(x[: x.numel() // 2]); a = x[: x.numel() // 2]
for _ in range(3): y.copy_(a)
torch.cuda.synchronize(); e0.record()
for _ in range(10): y.copy_(a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"copy (read+write): {2*a.numel()*4/ms/1e6:.1f} GB/s ({ms:.3f} ms)")
