"""Precision study (test infrastructure, not collected by pytest): how far does the decoded RGB-D move from the fp32
oracle when the operands of every conv / 1x1 / attention product are rounded as a candidate tensor-core scheme would?

    python tests/studies/split_precision_study.py [res]

Schemes (a = activation operand, w = weight operand; hi/lo = two-plane split of the fp32 value):
  bf16x3  a_hi*w_hi + a_hi*w_lo + a_lo*w_hi      (what the engine runs)            -> both operands ~16 mantissa bits
  f16x3   the same with fp16 planes                                                 -> both ~22 bits
  f16x2a  a_hi*(w_hi + w_lo): activations ONE fp16 plane, weights two               -> a: 11 bits
  f16x2w  (a_hi + a_lo)*w_hi: weights ONE fp16 plane, activations two               -> w: 11 bits
  bf16x2a / bf16x2w the same with bf16 planes                                       -> 8 bits
Only the decoder is perturbed (the encoder decides the tokens and stays on the 3-term scheme)."""
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import network, recipes  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def planes(x, dt, n):
    hi = x.to(dt).float()
    if n == 1:
        return hi
    return hi + (x - hi).to(dt).float()


SCHEMES = {
    "bf16x3": (torch.bfloat16, 2, 2), "f16x3": (torch.float16, 2, 2),
    "f16x2a": (torch.float16, 1, 2), "f16x2w": (torch.float16, 2, 1),
    "bf16x2a": (torch.bfloat16, 1, 2), "bf16x2w": (torch.bfloat16, 2, 1),
}


def run(sd, z, scheme):
    if scheme is None:
        return network.decoder(sd, z)
    dt, na, nw = SCHEMES[scheme]
    conv0, bmm0 = network._conv, torch.bmm

    def conv(sd_, name, x, stride=1, padding=0):
        return F.conv2d(planes(x, dt, na).double(), planes(sd_[f"{name}.weight"], dt, nw).double(), sd_[f"{name}.bias"].double(),
                        stride=stride, padding=padding).float()

    def bmm(a, b):                      # attention products: left operand plays the activation role, right the "weight" role
        return bmm0(planes(a, dt, na).double(), planes(b, dt, nw).double()).float()
    network._conv, torch.bmm = conv, bmm
    try:
        return network.decoder(sd, z)
    finally:
        network._conv, torch.bmm = conv0, bmm0


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    torch.set_grad_enabled(False)
    for ds in ("clevr-infinite", "google_earth"):
        n_e = recipes.DATASETS[ds]["n_embed"]
        sd = recipes.make_state_dict(n_e, seed=0)
        rng = np.random.default_rng(7)
        idx = rng.integers(0, n_e, size=(1, res // 16, res // 16))
        z = sd["quantize.embedding.weight"][torch.from_numpy(idx)].permute(0, 3, 1, 2).contiguous()
        z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
        ref = run(sd, z, None)
        for s in SCHEMES:
            out = run(sd, z, s)
            e = float((out - ref).norm() / ref.norm())
            em = float((out - ref).abs().max() / ref.abs().max())
            print(f"{ds:15s} res {res} {s:8s} rel-L2 {e:.3e}   max-abs/max {em:.3e}", flush=True)


if __name__ == "__main__":
    main()
