"""Dev probe (GPU box): which rounding does eager torch produce ON CUDA for the two ops that
decide the height-field cell indices (SURVEY.md §7 hard part 1)?"""
import numpy as np
import torch

torch.manual_seed(0)
n = 4_000_000
z, w = torch.randn(n), torch.randn(n)
q = torch.stack([torch.zeros(n), torch.zeros(n), z, w], -1).cuda()
nt = q.norm(p=2, dim=-1).cpu().numpy()
z32, w32 = z.numpy(), w.numpy()
z64, w64 = z32.astype(np.float64), w32.astype(np.float64)
var = {
    "sep rn(rn(z2)+rn(w2))": np.sqrt((z32 * z32 + w32 * w32).astype(np.float32)),
    "fma(w,w,rn(z2))": np.sqrt((w64 * w64 + (z32 * z32).astype(np.float64)).astype(np.float32)),
    "fma(z,z,rn(w2))": np.sqrt((z64 * z64 + (w32 * w32).astype(np.float64)).astype(np.float32)),
}
for k, v in var.items():
    print(f"norm variant {k:28s} mismatches vs torch.cuda: {(v != nt).sum()}")
x = (torch.rand(n) * 2300)
d = (x.cuda() / 0.1).cpu().numpy()
print("div: true-divide mismatches", (d != (x.numpy() / np.float32(0.1))).sum(), " mul-by-10 mismatches",
      (d != (x.numpy() * np.float32(10.0))).sum())
print(torch.__version__, torch.cuda.get_device_name(0))
