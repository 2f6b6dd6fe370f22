import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/img-compression-mps_b200")
from imgcompressionmps import _native, _ops
ctx = _native.context()
g = torch.Generator(device="cuda").manual_seed(3)
for n in (256, 512):
    a = torch.randn(n, 4 * n, dtype=torch.float64, device="cuda", generator=g)
    gm = a @ a.T
    for cl in (0, 1):
        ctx.set_option("topk_cluster", cl)
        print(f"n={n} cluster={cl}", file=sys.stderr, flush=True)
        for _ in range(2):
            _ops.eigh_topk(gm, 64)
        torch.cuda.synchronize()
