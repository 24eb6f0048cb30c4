"""CPU: precision of the split-fp16 operand scheme of the InfoNCE kernels (csrc/scp_nce.cu, split_f16 / nce_prep_kernel),
simulated exactly: x * kFeatScale = hi + lo with hi = fp16(x s), lo = fp16(x s - hi); the tensor cores accumulate
hi.hi + hi.lo + lo.hi in fp32 (operands laid out [hi|hi|lo] x [hi|lo|hi] along K).  The logits that enter exp() are
scaled by up to 1/0.07 (and by a learnable temperature that the reference clamps nowhere), so their absolute error has
to stay near fp32 round-off for the loss to be within the 1e-3 tolerance with room to spare; a plain fp16 product
(error ~1e-3 * scale) would not do.  The scale constant is parsed from the source."""
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "speechclip_plus_b200", "csrc", "scp_nce.cu")).read()
FEAT_SCALE = float(re.search(r"constexpr float kFeatScale = ([0-9.e+-]+)f?;", SRC).group(1))


def _split(x32: torch.Tensor):
    xs = x32 * FEAT_SCALE
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return hi.float(), lo.float()


def test_split_fp16_logits_are_fp32_accurate():
    g = torch.Generator().manual_seed(7122)
    for N, D in ((256, 512), (512, 768), (64, 64)):
        b = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
        a = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + 1.5 * b, dim=-1)
        ah, al = _split(a)
        bh, bl = _split(b)
        assert torch.isfinite(ah).all() and ah.abs().max() < 65504          # no fp16 overflow of the scaled operands
        # fp32 accumulation of the three partial products, then the common 1/scale^2
        s3 = (ah @ bh.t() + ah @ bl.t() + al @ bh.t()) / (FEAT_SCALE * FEAT_SCALE)
        exact = a.double() @ b.double().t()
        err3 = (s3.double() - exact).abs().max().item()
        err1 = ((a.half().float() @ b.half().float().t()).double() - exact).abs().max().item()
        assert err3 < 2e-6, (N, D, err3)          # ~fp32 round-off of a D-term dot product of unit vectors
        assert err1 > 20 * err3                    # what a single fp16 product would give
        # effect on the loss at the sharpest shipped temperature (scale 1/0.07): relative error of sum_j exp(S_ij)
        scale = 1.0 / 0.07
        z3 = torch.exp(s3.double() * scale).sum(1)
        z = torch.exp(exact * scale).sum(1)
        assert ((z3 - z).abs() / z).max().item() < 5e-5
