"""Diagnose first-call latencies (module load / table build / library init) of each op."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
t0 = time.time(); torch.zeros(1, device="cuda"); torch.cuda.synchronize(); print(f"cuda init {time.time()-t0:.2f}s")
from lightning_pose_b200 import ops
def tm(name, fn):
    for i in range(2):
        t0 = time.time(); fn(); torch.cuda.synchronize(); print(f"{name} call{i}: {time.time()-t0:.3f}s", flush=True)
dev = "cuda"
for ds in (1, 2, 3):
    x = torch.rand(1, 4, 8, 8, device=dev)
    tm(f"decode ds={ds} 8x8", lambda: ops.decode_softargmax(x, ds, 1000.0))
x = torch.rand(2, 17, 96, 96, device=dev)
tm("decode 96x96", lambda: ops.decode_softargmax(x, 2, 1000.0))
f = torch.randn(2, 64, 3, 4, device=dev); w1 = torch.randn(16, 5, 3, 3, device=dev); w2 = torch.randn(5, 5, 3, 3, device=dev); b = torch.zeros(5, device=dev)
tm("head f32", lambda: ops.head_forward(f, [w1, w2], [b, b]))
fr = f.clone().requires_grad_(True)
tm("torch convT bwd", lambda: torch.nn.functional.conv_transpose2d(torch.nn.functional.pixel_shuffle(fr, 2), w1, b, stride=2, padding=1, output_padding=1).sum().backward())
kp = torch.rand(2, 5, 2, device=dev) * 40
tm("generate_heatmaps", lambda: ops.generate_heatmaps(kp, 48, 64, (12, 16)))
