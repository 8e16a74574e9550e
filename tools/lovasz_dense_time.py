import sys, torch
sys.path.insert(0, "/root/repo")
from coarse3d_b200 import ops
for B, frac in ((8, 0.3), (8, 1.0), (64, 0.3)):
    C, H, W = 20, 64, 2048
    g = torch.Generator(device="cuda").manual_seed(1)
    probs = torch.softmax(torch.randn(B, C, H, W, device="cuda", generator=g), 1)
    labels = torch.randint(1, C, (B, H, W), device="cuda", generator=g) * (torch.rand(B, H, W, device="cuda", generator=g) < frac)
    n_valid = int((labels != 0).sum()); cap = 1024
    while cap < n_valid: cap *= 2
    cap = min(cap, labels.numel()) if n_valid <= 32768 else (n_valid + 4095) // 4096 * 4096
    p = probs.requires_grad_(True)
    for _ in range(2):
        p.grad = None
        loss, ws = ops.lovasz_softmax(p, labels, ignore=0, max_valid=cap); loss.backward()
    torch.cuda.synchronize()
    with ops.profile("") as prof:
        for _ in range(3):
            p.grad = None
            loss, ws = ops.lovasz_softmax(p, labels, ignore=0, max_valid=cap, workspace=ws); loss.backward()
        torch.cuda.synchronize()
        print(B, frac, n_valid, cap, float(loss), {k: round(1e3 * v[0] / max(v[1], 1), 1) for k, v in prof.all().items()})
    del p, probs, labels, ws
    torch.cuda.empty_cache()
