import torch, fimex_b200 as fb
g = torch.Generator(device="cuda").manual_seed(1)
shape = (4, 202, 1440)
f = torch.randn(shape, device="cuda", generator=g) + 280
f[torch.rand(shape, device="cuda", generator=g) < 0.05] = float("nan")
f[..., 50:90, 100:300] = float("nan")
for _ in range(2):
    d = f.clone(); fb.fill2d_device(d, 1e-9, 1.6, 100); torch.cuda.synchronize()
print("ok", int(torch.isnan(d).sum()))
