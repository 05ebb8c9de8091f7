"""The denoiser side of the mechanism call -- NOT part of the accelerated hot path (north-star: "the UNet forward stays
in PyTorch and is timed separately").

* `ADMUNet`: a plain-PyTorch guided-diffusion (ADM) UNet written from the architecture description, for timing the
  denoiser with random-init weights.  Flags of the two BASELINE configurations (upstream
  guided-diffusion / DPS -- the reference does not ship them, F14):
    imagenet256: model_channels 256, res blocks 2, attention at 32/16/8, head channels 64, channel mult (1,1,2,2,4,4),
                 scale-shift norm, resblock up/down, learn_sigma (6 output channels)
    ffhq256:     model_channels 128, res blocks 1, attention at 16, otherwise the same
  The reference's factory zero-initialises the last convolution (training/openai_unet.py:618-622), so a freshly built
  net outputs 0; `randomize_output=True` re-draws it so that timing runs see non-trivial activations.
* `IDDPMLinearPrecond`: the sigma-parametrised wrapper of training/openai_preconditioning.py:93-207 restated:
  c_in = 1/sqrt(sigma^2+1), c_out = -sigma, c_skip = 1, timestep = M - argmin_j |sigma - u_j|, D = clamp(x + c_out F, -1, 1)
  in the dtype promotion of the reference (float64 because c_out is float64).  Returns (D_x, None): the
  second output (x0_var) is only read by the Peng baselines.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .sampler import iddpm_linear_u

CONFIGS = {
    "imagenet256": dict(model_channels=256, num_res_blocks=2, attention_resolutions=(32, 16, 8), channel_mult=(1, 1, 2, 2, 4, 4)),
    "ffhq256": dict(model_channels=128, num_res_blocks=1, attention_resolutions=(16,), channel_mult=(1, 1, 2, 2, 4, 4)),
}


def _norm(ch):
    return nn.GroupNorm(32, ch)


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class ResBlock(nn.Module):
    """GroupNorm-SiLU-conv, scale-shift conditioning on the timestep embedding, optional 2x up / down sampling inside."""

    def __init__(self, cin, cout, emb_ch, up=False, down=False):
        super().__init__()
        self.up, self.down = up, down
        self.norm1, self.conv1 = _norm(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.emb = nn.Linear(emb_ch, 2 * cout)
        self.norm2, self.conv2 = _norm(cout), nn.Conv2d(cout, cout, 3, padding=1)
        nn.init.zeros_(self.conv2.weight); nn.init.zeros_(self.conv2.bias)
        self.skip = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)

    def _resample(self, x):
        if self.up:
            return F.interpolate(x, scale_factor=2, mode="nearest")
        if self.down:
            return F.avg_pool2d(x, 2)
        return x

    def forward(self, x, emb):
        h = F.silu(self.norm1(x))
        h, x = self._resample(h), self._resample(x)
        h = self.conv1(h)
        scale, shift = self.emb(F.silu(emb))[:, :, None, None].chunk(2, dim=1)
        h = self.norm2(h) * (1 + scale) + shift
        h = self.conv2(F.silu(h))
        return self.skip(x) + h


class AttentionBlock(nn.Module):
    def __init__(self, ch, head_ch=64):
        super().__init__()
        self.heads = ch // head_ch
        self.norm = _norm(ch)
        self.qkv = nn.Conv1d(ch, 3 * ch, 1)
        self.proj = nn.Conv1d(ch, ch, 1)
        nn.init.zeros_(self.proj.weight); nn.init.zeros_(self.proj.bias)

    def forward(self, x):
        b, c, h, w = x.shape
        qkv = self.qkv(self.norm(x).reshape(b, c, h * w)).reshape(b, 3, self.heads, c // self.heads, h * w)
        q, k, v = (t.transpose(-1, -2) for t in qkv.unbind(1))                  # (b, heads, tokens, head_ch)
        a = F.scaled_dot_product_attention(q, k, v)
        a = a.transpose(-1, -2).reshape(b, c, h * w)
        return x + self.proj(a).reshape(b, c, h, w)


class _Stage(nn.Module):
    def __init__(self, block, attn=None):
        super().__init__()
        self.block, self.attn = block, attn

    def forward(self, x, emb):
        x = self.block(x, emb) if isinstance(self.block, ResBlock) else self.block(x)
        return self.attn(x) if self.attn is not None else x


class ADMUNet(nn.Module):
    def __init__(self, image_size=256, in_channels=3, out_channels=6, model_channels=256, num_res_blocks=2,
                 attention_resolutions=(32, 16, 8), channel_mult=(1, 1, 2, 2, 4, 4), head_channels=64, randomize_output=True):
        super().__init__()
        mc, emb_ch = model_channels, 4 * model_channels
        self.model_channels = mc
        self.time_embed = nn.Sequential(nn.Linear(mc, emb_ch), nn.SiLU(), nn.Linear(emb_ch, emb_ch))
        self.down = nn.ModuleList([_Stage(nn.Conv2d(in_channels, mc, 3, padding=1))])
        chans, ch, res = [mc], mc, image_size
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                blk = ResBlock(ch, mult * mc, emb_ch)
                ch = mult * mc
                self.down.append(_Stage(blk, AttentionBlock(ch, head_channels) if res in attention_resolutions else None))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.down.append(_Stage(ResBlock(ch, ch, emb_ch, down=True)))
                chans.append(ch)
                res //= 2
        self.mid = nn.ModuleList([_Stage(ResBlock(ch, ch, emb_ch), AttentionBlock(ch, head_channels)), _Stage(ResBlock(ch, ch, emb_ch))])
        self.up = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                blk = ResBlock(ch + chans.pop(), mult * mc, emb_ch)
                ch = mult * mc
                stage = [_Stage(blk, AttentionBlock(ch, head_channels) if res in attention_resolutions else None)]
                if level and i == num_res_blocks:
                    stage.append(_Stage(ResBlock(ch, ch, emb_ch, up=True)))
                    res *= 2
                self.up.append(nn.ModuleList(stage))
        self.out_norm = _norm(ch)
        self.out_conv = nn.Conv2d(ch, out_channels, 3, padding=1)
        if not randomize_output:
            nn.init.zeros_(self.out_conv.weight); nn.init.zeros_(self.out_conv.bias)

    def load_reference_state_dict(self, sd, legacy_attention_order=True):
        """Loads a state dict of the reference's UNetModel (training/openai_unet.py:395-686; the published ADM / DPS
        checkpoints have this layout).  Key map: input_blocks.N.{0,1} -> down[N].{block,attn}; middle_block.{0,1,2} ->
        mid[0].block, mid[0].attn, mid[1].block; output_blocks.N.j -> the j-th module of up[N] in the order ResBlock,
        attention, up-sampling ResBlock; inside a ResBlock in_layers.{0,2} / emb_layers.1 / out_layers.{0,3} /
        skip_connection -> norm1, conv1 / emb / norm2, conv2 / skip; proj_out -> proj; out.{0,2} -> out_norm, out_conv.
        The reference's default QKVAttentionLegacy (:327-353) lays the 3C rows of `qkv` out as [head][q,k,v][channel];
        this class uses [q,k,v][head][channel], so the rows are permuted on the way in."""
        res_map = {"in_layers.0": "norm1", "in_layers.2": "conv1", "emb_layers.1": "emb", "out_layers.0": "norm2",
                   "out_layers.3": "conv2", "skip_connection": "skip"}

        def sub(prefix):
            return {k[len(prefix):]: torch.as_tensor(v) for k, v in sd.items() if k.startswith(prefix)}

        def load_module(mod, prefix):
            part = sub(prefix)
            if isinstance(mod, ResBlock):
                out = {}
                for k, v in part.items():
                    head, leaf = k.rsplit(".", 1)
                    out[f"{res_map[head]}.{leaf}"] = v
                part = out
            elif isinstance(mod, AttentionBlock):
                out = {}
                for k, v in part.items():
                    k = k.replace("proj_out.", "proj.")
                    if k.startswith("qkv.") and legacy_attention_order:
                        c = v.shape[0] // 3
                        v = v.reshape(mod.heads, 3, c // mod.heads, *v.shape[1:]).transpose(0, 1).reshape(v.shape)
                    out[k] = v
                part = out
            mod.load_state_dict(part, strict=True)
            return len(part)

        used = load_module(self.time_embed, "time_embed.")
        for n, st in enumerate(self.down):
            used += load_module(st.block, f"input_blocks.{n}.0.")
            if st.attn is not None:
                used += load_module(st.attn, f"input_blocks.{n}.1.")
        used += load_module(self.mid[0].block, "middle_block.0.")
        used += load_module(self.mid[0].attn, "middle_block.1.")
        used += load_module(self.mid[1].block, "middle_block.2.")
        for n, stages in enumerate(self.up):
            j = 0
            for st in stages:
                used += load_module(st.block, f"output_blocks.{n}.{j}.")
                j += 1
                if st.attn is not None:
                    used += load_module(st.attn, f"output_blocks.{n}.{j}.")
                    j += 1
        used += load_module(self.out_norm, "out.0.")
        used += load_module(self.out_conv, "out.2.")
        if used != len(sd):
            raise KeyError(f"reference state dict has {len(sd)} entries, {used} were consumed")
        return self

    def forward(self, x, timesteps):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels).to(x.dtype))
        hs, h = [], x
        for st in self.down:
            h = st(h, emb)
            hs.append(h)
        for st in self.mid:
            h = st(h, emb)
        for stages in self.up:
            h = torch.cat([h, hs.pop()], dim=1)
            for st in stages:
                h = st(h, emb)
        return self.out_conv(F.silu(self.out_norm(h)))


class IDDPMLinearPrecond(nn.Module):
    def __init__(self, net, img_channels=3, M=1000, use_fp16=True):
        super().__init__()
        self.net, self.img_channels, self.M, self.use_fp16 = net, img_channels, M, use_fp16
        self.register_buffer("u", iddpm_linear_u(M))

    def round_sigma_index(self, sigma):
        return torch.cdist(sigma.to(torch.float32).reshape(1, -1, 1), self.u.reshape(1, -1, 1)).argmin(2).flatten()

    def forward(self, x, sigma):
        sigma = torch.as_tensor(sigma, device=x.device).to(torch.float64).reshape(-1, 1, 1, 1)
        x32 = x.to(torch.float32)
        c_in = 1 / (sigma ** 2 + 1).sqrt()
        t = (self.M - self.round_sigma_index(sigma)).to(torch.long).repeat(x.shape[0] // sigma.shape[0])
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.use_fp16 and x.is_cuda):
            out = self.net(c_in.to(torch.float32) * x32, t)
        f_x = out[:, : self.img_channels].to(torch.float32)
        d_x = x32 + (-sigma) * f_x                         # float64 through c_out, like the reference
        return torch.clamp(d_x, -1, 1), None


def build_denoiser(config="imagenet256", device="cuda", use_fp16=True):
    net = ADMUNet(**CONFIGS[config]).to(device)
    return IDDPMLinearPrecond(net, use_fp16=use_fp16).to(device).eval()


class GraphedDenoiser:
    """CUDA-graph capture of the denoiser's forward AND its VJP: `torch.cuda.make_graphed_callables`
    records the forward pass and the backward pass of `den(x, sigma)` once; every later mechanism call replays the two
    graphs (the VJP when the mechanism calls `autograd.grad(..., x_t)`, conditioning_mechanisms.py:280) instead of
    issuing the UNet's ~1500 kernels from Python.  sigma travels through a static device scalar, so one capture serves
    every noise level; x must keep the captured shape.  Returns (D_x, None) like the wrapper it graphs."""

    def __init__(self, den, sample_x):
        self.den = den
        self.sig = torch.zeros((), dtype=torch.float64, device=sample_x.device)

        class _Fn(nn.Module):
            def __init__(self, d):
                super().__init__()
                self.d = d

            def forward(self, x, sigma):
                return self.d(x, sigma)[0]
        self.fn = _Fn(den)
        x = sample_x.detach().clone().requires_grad_()
        self.sig.fill_(1.0)
        self.graphed = torch.cuda.make_graphed_callables(self.fn, (x, self.sig))

    def __call__(self, x, sigma):
        self.sig.fill_(float(sigma))
        # the graph writes its output into a static buffer that the next replay overwrites: the mechanism keeps the
        # previous call's mean as history, so it gets its own copy (clone is differentiable: the VJP still replays)
        return self.graphed(x, self.sig).clone(), None
