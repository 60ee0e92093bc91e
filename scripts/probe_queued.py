import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from lightning_pose_b200 import ops
from lightning_pose_b200.models.heads.heatmap import HeatmapHead
dev = torch.device("cuda:0")
prob = bench.make_problem(4, seed=1234, device=dev, regime="trained")
head = HeatmapHead("resnet50", 2048, 17)
d1, d2 = list(head.upsampling_layers)[1:]
with torch.no_grad():
    w1, b1, w2, b2 = prob["head_params"]
    d1.weight.copy_(w1), d1.bias.copy_(b1), d2.weight.copy_(w2), d2.bias.copy_(b2)
head = head.to(dev)
with torch.no_grad():
    hm = head(prob["feats"].bfloat16().to(dev))
    xy, conf, stats = ops._decode_fwd(hm, 2, 1000.0)
st = stats.reshape(-1, 8)
rows = st[:, 5] - st[:, 4] + 1
cols = st[:, 7] - st[:, 6] + 1
big = (rows > 22) | (cols > 22)
print("planes", st.shape[0], "box too big for the warp window:", int(big.sum()))
hmf = hm.reshape(-1, 96, 96)
idx = torch.nonzero(big).flatten()[:12]
for i in idx.tolist():
    p = hmf[i]
    mx = float(p.max()); second = float(p.flatten().topk(40).values[-1])
    am = int(p.argmax()); 
    print(i, "frame", i // 17, "kp", i % 17, "max %.4f 40th %.5f argmax (%d,%d) box rows %d cols %d conf %.3f" % (mx, second, am // 96, am % 96, int(rows[i]), int(cols[i]), float(conf.flatten()[i])))
print("hist rows:", torch.bincount(rows.clamp(0, 40).long()).tolist())
