import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lightning_pose_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
n, k, s = (int(x) for x in sys.argv[1:4])
hm = torch.softmax(torch.randn(n, k, s * s, device=dev) * 2.0, -1).reshape(n, k, s, s)
kp, cf = ops.decode_softargmax(hm, 2, 1000.0)
torch.cuda.synchronize()
print("ok", n, k, s, float(kp.mean()), float(cf.mean()), flush=True)
