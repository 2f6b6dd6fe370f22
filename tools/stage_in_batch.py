"""Stage times per volume (CUDA events of the library's stage profiler) with several volumes in flight:
shows which stages stretch when kernels of different volumes share the GPU."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "img-compression-mps_b200"))
sys.path.insert(0, str(ROOT))
from bench import synthetic_volume  # noqa: E402
from imgcompressionmps.batch import VolumePipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nvol = int(sys.argv[2]) if len(sys.argv) > 2 else 48
torch.cuda.set_device(0)
vols = [torch.from_numpy(synthetic_volume((n, n, n), 2026 + i)).cuda() for i in range(4)]
items = [vols[i % 4] for i in range(nvol)]
for w in [int(v) for v in sys.argv[3:]] or [1, 4, 12]:
    with VolumePipeline(workers=w) as pipe:
        pipe.roundtrip(items[:2 * w], max_bond=64, keep=False)
        pipe.roundtrip(items[:2 * w], max_bond=64, keep=False)
        for c in pipe._contexts:
            c.profile(True)
            c.stage_times(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.roundtrip(items, max_bond=64, keep=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tot = {}
        for c in pipe._contexts:
            for k, (ms, calls) in c.stage_times().items():
                tot[k] = tot.get(k, 0.0) + ms
            c.profile(False)
        print(f"{n}^3, {w} in flight: {dt * 1e3 / nvol:.2f} ms/volume; stage ms per volume:",
              {k: round(v / nvol, 3) for k, v in tot.items() if v > 0}, flush=True)
