import sys, torch
sys.path.insert(0, "img-compression-mps_b200"); sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from imgcompressionmps import _ops
from perf_probe import timeit
g = torch.Generator(device="cuda").manual_seed(0)
for shape in ((1920, 1080, 64), (256, 256, 256), (64, 64, 32, 400)):
    x = torch.rand(shape, dtype=torch.float32, device="cuda", generator=g)
    ms = timeit(lambda: _ops.dct_last_axis(x), reps=3, warm=1)
    ms2 = timeit(lambda: _ops.dct_last_axis(x, inverse=True), reps=3, warm=1)
    n = shape[-1]
    print(f"dct {shape}: forward {ms:.3f} ms, inverse {ms2:.3f} ms ({x.numel() * 8 / ms / 1e6:.0f} GB/s algorithmic, {2 * n * x.numel() / ms / 1e9:.2f} TFLOP/s)", flush=True)
