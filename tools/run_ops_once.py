"""Launches every standalone geometry op once at the BASELINE.json configs[4] size ([8,3,768,1024]) -- for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demon_b200 import lmbspecialops as ops

n, h, w = 8, 768, 1024
img = torch.rand(n, 3, h, w, device="cuda") - 0.5
yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h, device="cuda"), torch.linspace(0, 6.28, w, device="cuda"), indexing="ij")
disp = (0.025 * torch.stack([torch.sin(xx + 0.5 * yy), torch.cos(yy - 0.3 * xx)])[None].repeat(n, 1, 1, 1)).contiguous()
depth = torch.rand(n, 1, h, w, device="cuda") + 0.3
K = torch.tensor([[0.89115971, 1.18821287, 0.5, 0.5]], device="cuda").repeat(n, 1)
r = (torch.rand(n, 3, device="cuda") - 0.5) * 0.1
t = torch.tensor([[0.9, 0.1, -0.05]], device="cuda").repeat(n, 1)
torch.cuda.synchronize()
ops.warp2d(img, disp, normalized=True, border_mode="value")
ops.depth_to_flow(depth, K, r, t, inverse_depth=True, normalize_flow=True)
ops.flow_to_depth(disp, K, r, t, normalized_flow=True, inverse_depth=True, nowarning=True)
ops.median3x3_downsample(img)
ops.scale_invariant_gradient(depth, [1, 2, 4, 8, 16], [1, .5, .25, .125, .0625])
ops.leaky_relu(img)
torch.cuda.synchronize()
print("ops done")
