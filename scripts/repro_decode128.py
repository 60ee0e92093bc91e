import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lightning_pose_b200 import ops
from lightning_pose_b200._lib import lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1]
n, k, s = (int(x) for x in sys.argv[2:5])
if which == "flat":
    hm = torch.softmax(torch.randn(n, k, s * s, device=dev) * 2.0, -1).reshape(n, k, s, s)
else:
    yy, xx = torch.meshgrid(torch.arange(s, device=dev), torch.arange(s, device=dev), indexing="ij")
    c = torch.rand(n, k, 2, device=dev) * (s - 20) + 10
    hm = torch.exp(-((yy[None, None] - c[..., 1, None, None]) ** 2 + (xx[None, None] - c[..., 0, None, None]) ** 2) / (2 * 1.3**2))
    hm = hm / hm.sum((2, 3), keepdim=True)
for ring in (0, 1):
    lib.lpb_set_tuning(3, ring)
    kp, cf = ops.decode_softargmax(hm, 2, 1000.0)
    torch.cuda.synchronize()
    print(which, n, k, s, "ring", ring, "ok", float(kp.mean()), float(cf.mean()), flush=True)
