"""B=768 head forward(train)+backward once or twice (for ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
from lightning_pose_b200 import ops  # noqa: E402

B = 768
feat = torch.randn(B, 2048, 12, 12, device="cuda").bfloat16()
w1 = torch.randn(512, 17, 3, 3, device="cuda") * 0.05
w2 = torch.randn(17, 17, 3, 3, device="cuda") * 0.2
b = torch.zeros(17, device="cuda")
gl = torch.randn(B, 17, 96, 96, device="cuda")
for _ in range(2):
    out, saved = ops._head_forward_bf16(feat, [w1, w2], [b, b], False, train=True)
    ops.head_backward_bf16(gl, saved, feat.shape, w1, w2)
torch.cuda.synchronize()
print("done")
