"""Flag / bbox statistics of the decode windows on the bench's trained-like heatmaps."""
import sys
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from lightning_pose_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
prob = bench.make_problem(16, seed=1234, device=dev, regime="trained")
head = prob["head"].to(dev)
feats = prob["feats"].to(torch.bfloat16).to(dev)
with torch.no_grad():
    hm = head(feats[256:])
xy, conf, stats = ops._decode_fwd(hm, 2, 1000.0)
st = stats.reshape(-1, 8)
nr = (st[:, 5] - st[:, 4] + 1)
nc = (st[:, 7] - st[:, 6] + 1)
print("bbox rows: mean %.1f max %d  >24: %d ; cols: mean %.1f max %d >24: %d of %d" % (nr.mean(), nr.max(), (nr > 24).sum(), nc.mean(), nc.max(), (nc > 24).sum(), nr.numel()))
print("hist rows", torch.bincount(nr.long().clamp(max=40)).tolist())
gxy = torch.randn(hm.shape[0], hm.shape[1], 2, device=dev)
for _ in range(3):
    win, meta, ov = ops.decode_backward_windows(hm, stats, gxy, 2, 1000.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    win, meta, ov = ops.decode_backward_windows(hm, stats, gxy, 2, 1000.0)
e1.record()
torch.cuda.synchronize()
print("windows call %.1f us" % (e0.elapsed_time(e1) * 100))
print("flags", torch.bincount(meta[:, 2].long()).tolist())
