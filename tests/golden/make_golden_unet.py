"""Golden vectors for the denoiser side: the REAL reference's UNetModel (training/openai_unet.py:395-686)
and its iDDPMLinearPrecond wrapper (training/openai_preconditioning.py:93-207) at a small configuration with every
structural feature of the BASELINE nets (scale-shift norm, residual up/down blocks, legacy-order multi-head attention on
two levels and in the middle block, learn_sigma's 6 output channels).  The zero-initialised layers are re-drawn so the
output is not 0 and rounded to bfloat16-representable values (the file keeps their upper 16 bits).  Stores the reference's state dict, the inputs and its outputs:  python tests/golden/make_golden_unet.py
Run in the build container only (the reference is not available on the GPU box).  Nothing here is product code."""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
CFG = dict(image_size=32, in_channels=3, model_channels=32, out_channels=6, num_res_blocks=2, attention_resolutions=(2, 4),
           channel_mult=(1, 1, 2), num_head_channels=16, use_scale_shift_norm=True, resblock_updown=True)

if __name__ == "__main__":
    sys.path.insert(0, REF)
    from training.openai_unet import UNetModel
    from training.openai_preconditioning import iDDPMLinearPrecond
    torch.manual_seed(0)
    net = UNetModel(**CFG, use_fp16=False, use_new_attention_order=False).eval()
    with torch.no_grad():
        for p in net.parameters():
            if float(p.abs().max()) == 0.0:
                p.normal_(0, 0.05)
            p.copy_(p.to(torch.bfloat16).to(torch.float32))       # bfloat16-representable weights: stored as 16-bit patterns
    g = {"sd/" + k: (v.view(torch.int32) >> 16).to(torch.int16).numpy() for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, 32, 32)
    t = torch.tensor([17, 803])
    with torch.no_grad():
        g["x"], g["t"], g["y"] = x.numpy(), t.numpy(), net(x, t).numpy()
    pre = iDDPMLinearPrecond(net, 32, 3, use_fp16=False).eval()
    g["u"] = pre.u.numpy()
    sig = [80.0, 2.37, 0.0413]
    for i, s in enumerate(sig):
        xs = (0.5 * torch.randn(2, 3, 32, 32) + s * torch.randn(2, 3, 32, 32)).to(torch.float64)
        with torch.no_grad():
            d, _ = pre(xs, torch.tensor(s))
        g[f"pre{i}_x"], g[f"pre{i}_sigma"], g[f"pre{i}_d"] = xs.numpy(), np.float64(s), d.numpy()
        g[f"pre{i}_dtype"] = str(d.dtype)
        g[f"pre{i}_index"] = pre.round_sigma(torch.tensor(s), return_index=True).numpy()
    np.savez_compressed(os.path.join(OUT, "unet_small.npz"), **g)
    print("unet_small.npz:", len(g), "arrays,", sum(v.size for v in g.values() if hasattr(v, "size")), "values")
