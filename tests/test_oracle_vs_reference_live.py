"""CPU, build container only (skipped where /root/reference is absent): the oracle against the REFERENCE'S OWN MODULES on
randomised inputs -- shapes, flags and seeds drawn per case -- for every component of the path.  The committed fixtures
(tests/golden/*.npz) pin the oracle at 37 hand-picked points; this sweep pins it on a few hundred more, including the
option combinations no fixture holds (margin x dcl x direction switches of the loss, every batch-norm layout x eval mode,
ragged CIF inputs with multiple fires per source, dynamic keyword counts at the edges 0 and Kmax).

Reference entry points (paths under /root/reference): avssl/module/weighted_sum.py, avssl/module/losses.py,
avssl/module/speechclip_c_modules/{my_vector_quantizer,kw_bn}.py, avssl/module/cif.py, avssl/model/kw_branches.py:158-197,
avssl/module/clip_official.py:222-279, avssl/util/data_utils.py:6-22.
"""
import math
import os
import sys
import types

import pytest
import torch

from conftest import norm_err, rel_err
from oracle import speechclip_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _ref_import as ref  # noqa: E402

pytestmark = pytest.mark.skipif(not ref.reference_available(), reason="needs the reference checkout (/root/reference)")

N_CASES = 24


def _gen(seed):
    return torch.Generator().manual_seed(7122 + seed)


def _randint(g, lo, hi):
    return int(torch.randint(lo, hi + 1, (1,), generator=g))


def _coin(g):
    return bool(torch.randint(0, 2, (1,), generator=g))


@pytest.fixture(scope="module")
def refmods():
    return types.SimpleNamespace(
        wsum=ref.load_leaf("avssl/module/weighted_sum.py", "live_ref_wsum"),
        losses=ref.load_leaf("avssl/module/losses.py", "live_ref_losses"),
        vq=ref.load_leaf("avssl/module/speechclip_c_modules/my_vector_quantizer.py", "live_ref_vq"),
        bn=ref.load_leaf("avssl/module/speechclip_c_modules/kw_bn.py", "live_ref_bn"),
        cif=ref.load_leaf("avssl/module/cif.py", "live_ref_cif"))


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(N_CASES))
def test_weighted_sum_live(refmods, seed):
    g = _gen(seed)
    L, B, T, D = _randint(g, 1, 25), _randint(g, 1, 5), _randint(g, 1, 40), 4 * _randint(g, 1, 48)
    norm = _coin(g)
    layer = refmods.wsum.WeightedSumLayer(n_weights=L, normalize_features=norm)
    with torch.no_grad():
        layer.weights.copy_(torch.randn(L, generator=g) * (0.0 if seed == 0 else 0.7))   # seed 0: the zero init
    storage = [torch.randn(T, B, D, generator=g) * (1 + l) + 0.1 * l for l in range(L)]
    layers = [s.transpose(0, 1).requires_grad_(True) for s in storage]                   # (T,B,D) storage, as the caller's
    gy = torch.randn(B, T, D, generator=g)
    y = layer(layers)
    grads = torch.autograd.grad(y, [layer.weights] + layers, grad_outputs=gy)
    w = layer.weights.detach().clone().requires_grad_(True)
    mine = [l.detach().clone().requires_grad_(True) for l in layers]
    y2 = oracle.wsum_forward(mine, w, normalize_features=norm)
    assert rel_err(y2, y) < 1e-5
    g2 = torch.autograd.grad(y2, [w] + mine, grad_outputs=gy)
    assert rel_err(g2[0], grads[0]) < 1e-4
    assert rel_err(oracle.wsum_grad_weights([l.detach() for l in layers], w.detach(), gy, normalize_features=norm),
                   grads[0]) < 1e-4
    for a, b in zip(g2[1:], grads[1:]):
        assert norm_err(a, b) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(N_CASES))
def test_masked_contrastive_loss_live(refmods, seed):
    g = _gen(100 + seed)
    N, D = _randint(g, 2, 300), 8 * _randint(g, 1, 16)
    trainable, dcl = _coin(g), _coin(g)
    margin = 0.0 if _coin(g) else float(torch.rand(1, generator=g)) * 0.3
    a2b, b2a = [(True, True), (True, False), (False, True)][_randint(g, 0, 2)]
    temperature = [0.07, 0.1, 0.5][_randint(g, 0, 2)]
    refmods.losses.MAX_EYE = max(256, N)                    # the reference indexes 256x256 buffers (losses.py:126)
    crit = refmods.losses.MaskedContrastiveLoss(temperature=temperature, temperature_trainable=trainable, margin=margin,
                                                dcl=dcl, a2b=a2b, b2a=b2a)
    b = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1).requires_grad_(True)
    a = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + 1.5 * b.detach(), dim=-1).requires_grad_(True)
    ids = [None, torch.arange(N), torch.randint(0, max(N // 5, 1), (N,), generator=g)][_randint(g, 0, 2)]
    loss = crit(a, b, ids)
    params = [a, b] + ([crit.temperature] if trainable else [])
    grads = torch.autograd.grad(loss, params)
    t = torch.tensor(math.log(1.0 / temperature), requires_grad=True)
    scale = t.exp() if trainable else 1.0 / temperature
    a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    kw = dict(margin=margin, dcl=dcl, a2b=a2b, b2a=b2a)
    loss2 = oracle.nce_forward(a2, b2, ids, scale, **kw)
    assert rel_err(loss2, loss) < 2e-5, (N, D, kw, trainable)
    g2 = torch.autograd.grad(loss2, [a2, b2] + ([t] if trainable else []))
    assert norm_err(g2[0], grads[0]) < 1e-4 and norm_err(g2[1], grads[1]) < 1e-4
    da, db, dlog = oracle.nce_grads(a.detach(), b.detach(), ids, float(torch.as_tensor(scale).detach()), **kw)          # the closed form the kernel uses
    assert norm_err(da, grads[0]) < 1e-4 and norm_err(db, grads[1]) < 1e-4
    if trainable:
        assert rel_err(g2[2], grads[2]) < 1e-4 and rel_err(dlog, grads[2]) < 5e-4


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(N_CASES))
def test_vector_quantizer_live(refmods, seed):
    ref.import_avssl()
    import avssl.model.kw_branches as kb
    g = _gen(200 + seed)
    B, K, V, D = _randint(g, 2, 6), _randint(g, 1, 12), _randint(g, 8, 900), 8 * _randint(g, 1, 12)
    training = seed % 3 != 0
    tau = [0.1, 0.07, 0.5, 1.0][_randint(g, 0, 3)]
    table = torch.randn(V, D, generator=g) * 0.02 + 0.003 * torch.randn(1, D, generator=g)
    kw = (torch.randn(B, K, D, generator=g) * table.std(0) + table.mean(0)).requires_grad_(True)
    branch = kb.GeneralBranch.__new__(kb.GeneralBranch)
    torch.nn.Module.__init__(branch)
    emb = torch.nn.Embedding(V, D)
    with torch.no_grad():
        emb.weight.copy_(table)
    emb.weight.requires_grad_(False)
    branch.text_dim = D
    branch.clip = types.SimpleNamespace(model=types.SimpleNamespace(token_embedding=emb))
    branch.linear_proj = torch.nn.Identity()
    hard = seed % 4 != 1                                             # every fourth case: hard=False (:130-131 without the STE term)
    branch.vector_quantizer = refmods.vq.SimpleVectorQuantizer(temp=f"fixed={tau}", hard=hard)
    branch.train(training)
    cos = branch.get_keyword_cosine_score(kw.detach())
    res, out = branch.vq_audio_features(kw)
    assert rel_err(oracle.cosine_scores_loop(kw.detach(), table), cos) < 1e-6
    assert rel_err(oracle.cosine_scores(kw.detach(), table), cos) < 1e-5
    kw2 = kw.detach().clone().requires_grad_(True)
    res2, out2 = oracle.vq_audio_features(kw2, table, torch.tensor([tau]), training=training, faithful_loop=True, hard=hard)
    assert torch.equal(res2["targets"], res["targets"])
    assert rel_err(res2["subword_prob"], res["subword_prob"]) < 1e-5
    assert rel_err(out2, out) < 1e-5
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t", "diversity_loss"):
        assert rel_err(res2[key], res[key]) < 2e-5, key
    assert res2["num_vars"] == res["num_vars"] and math.isclose(res2["temp"], res["temp"], rel_tol=1e-6)
    if training:
        gout = torch.randn(B, K, D, generator=g)
        (gr,) = torch.autograd.grad(out, [kw], grad_outputs=gout)
        (g2,) = torch.autograd.grad(out2, [kw2], grad_outputs=gout)
        assert norm_err(g2, gr) < 1e-4
        closed, _ = oracle.vq_keyword_grad(kw.detach(), table, torch.tensor([tau]), gout)   # the form the kernel implements
        assert norm_err(closed, gr) < 2e-4
    else:
        assert not out.requires_grad and not out2.requires_grad


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(N_CASES))
def test_kw_batchnorm_live(refmods, seed):
    g = _gen(300 + seed)
    B, K, D = _randint(g, 2, 9), _randint(g, 1, 10), 4 * _randint(g, 1, 24)
    kind = ["eachKw_parallel", "eachKw_layers", "same", "same_seqlens", "dynamic"][seed % 5]
    training = seed % 4 != 3
    table = torch.randn(200, D, generator=g) * 0.02 + 0.003 * torch.randn(1, D, generator=g)
    init_bias, init_scale = table.mean(0), table.std(0)
    std_scale = [1, 2.5][_randint(g, 0, 1)]
    if kind == "dynamic":
        layer = refmods.bn.Kw_BatchNorm_dynamic(kw_dim=D, init_bias=init_bias, init_scale=init_scale, std_scale=std_scale)
    else:
        layer = refmods.bn.Kw_BatchNorm(kw_num=K, kw_dim=D, batchnorm_type="eachKw" if kind.startswith("eachKw") else "same",
                                        init_bias=init_bias, init_scale=init_scale, std_scale=std_scale, learnable=True,
                                        parallel=(kind == "eachKw_parallel"))
    bns = list(layer.bn_layers) if hasattr(layer, "bn_layers") else [layer.bn_layer]
    for bn in bns:
        with torch.no_grad():
            bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.1)
            bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
    stacked = kind == "eachKw_layers"
    pick = (lambda ts: torch.stack(ts)) if stacked else (lambda ts: ts[0])
    w0 = pick([bn.weight.detach().clone() for bn in bns]).requires_grad_(True)
    b0 = pick([bn.bias.detach().clone() for bn in bns]).requires_grad_(True)
    rm0 = pick([bn.running_mean.clone() for bn in bns])
    rv0 = pick([bn.running_var.clone() for bn in bns])
    layer.train(training)
    seq_lens = None
    if kind == "same_seqlens":
        seq_lens = [_randint(g, 1, K) for _ in range(B)]
        seq_lens[0] = K
    x = (torch.randn(B, K, D, generator=g) * 0.7 + 0.3 * torch.randn(1, 1, D, generator=g)).requires_grad_(True)
    x_in = x.clone()                                             # the seq_lens branch writes into its input (kw_bn.py:157)
    y = layer(x_in, torch.tensor(seq_lens)) if seq_lens is not None else layer(x_in)
    gy = torch.randn(B, K, D, generator=g)
    params = [p for bn in bns for p in (bn.weight, bn.bias)]
    grads = torch.autograd.grad(y, [x] + params, grad_outputs=gy)
    x2 = x.detach().clone().requires_grad_(True)
    y2, rm, rv = oracle.kw_batchnorm(x2, w0, b0, rm0, rv0, "eachKw" if kind.startswith("eachKw") else "same",
                                     parallel=(kind == "eachKw_parallel"), training=training, seq_lens=seq_lens)
    assert rel_err(y2, y) < 2e-5, kind
    assert rel_err(rm, pick([bn.running_mean for bn in bns])) < 2e-5
    assert rel_err(rv, pick([bn.running_var for bn in bns])) < 2e-5
    gx, gw, gb = torch.autograd.grad(y2, [x2, w0, b0], grad_outputs=gy)
    assert rel_err(gx, grads[0]) < 2e-4
    assert rel_err(gw, pick(list(grads[1::2]))) < 2e-4 and rel_err(gb, pick(list(grads[2::2]))) < 2e-4


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(N_CASES))
def test_cif_integrate_and_fire_live(refmods, seed):
    g = _gen(400 + seed)
    B, S, C = _randint(g, 1, 6), _randint(g, 4, 60), 4 * _randint(g, 1, 16)
    mode = ["train", "train_multifire", "infer", "notail"][seed % 4]
    x = torch.randn(B, S, C, generator=g).requires_grad_(True)
    a_scale = 0.5 if mode == "train_multifire" else float(torch.rand(1, generator=g)) * 0.3 + 0.15
    raw = torch.rand(B, S, generator=g) * 2 * a_scale
    lens = torch.randint(max(S // 2, 1), S + 1, (B,), generator=g)
    lens[0] = S
    raw = raw.masked_fill(torch.arange(S)[None, :] >= lens[:, None], 0.0).requires_grad_(True)
    layer = refmods.cif.CIF(cif_threshold=1.0, cif_output_dim=C, encoder_embed_dim=C, apply_tail_handling=(mode != "notail"))
    target, alpha = None, raw
    if mode.startswith("train"):
        hi = min(2 * S, 70) if mode == "train_multifire" else max(S // 3, 1)     # more targets than sources: several fires
        target = torch.randint(1, hi + 1, (B,), generator=g)
        alpha = raw * ((1.0 * target.type_as(raw) + 1e-5) / raw.sum(1)).unsqueeze(1)          # cif.py:126-129
    out = layer.integrate_and_fire(x, alpha, target_lengths=target)
    x2, raw2 = x.detach().clone().requires_grad_(True), raw.detach().clone().requires_grad_(True)
    alpha2 = raw2 if target is None else raw2 * ((1.0 * target.type_as(raw2) + 1e-5) / raw2.sum(1)).unsqueeze(1)
    feats, feat_len, fired = oracle.cif_integrate_and_fire(x2, alpha2, 1.0, target, mode != "notail")
    assert torch.equal(feat_len, out["dsample_feats_length"]), mode
    assert torch.equal(fired, out["fired_marks"])
    assert feats.shape == out["dsample_feats"].shape and rel_err(feats, out["dsample_feats"]) < 2e-5
    assert torch.equal(oracle.keypadding_mask(feats.shape[1], feat_len), out["dsample_feats_pad_mask"])
    if mode != "infer":   # the reference's inference tail path cannot be back-propagated (in-place edit of a saved tensor)
        gy = torch.randn(feats.shape, generator=g)
        gx, ga = torch.autograd.grad(out["dsample_feats"], [x, raw], grad_outputs=gy)
        gx2, ga2 = torch.autograd.grad(feats, [x2, raw2], grad_outputs=gy)
        assert rel_err(gx2, gx) < 1e-4 and rel_err(ga2, ga) < 2e-4


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(N_CASES))
def test_keyword_splice_live(seed):
    ref.import_avssl()
    import avssl.module.clip_official as co
    import avssl.util.data_utils as du
    g = _gen(500 + seed)
    B, Kmax, D, V, L = _randint(g, 1, 6), _randint(g, 1, 20), 8 * _randint(g, 1, 8), _randint(g, 10, 300), 77
    dynamic = seed % 2 == 1
    emb = torch.nn.Embedding(V, D)
    with torch.no_grad():
        emb.weight.copy_(torch.randn(V, D, generator=g) * 0.02)
    emb.weight.requires_grad_(False)
    captured = {}

    class Tower(torch.nn.Module):
        def forward(self, x):                       # (L, N, D)
            captured["x"] = x.permute(1, 0, 2)
            return x * 1.5

    model = types.SimpleNamespace(token_embedding=emb, positional_embedding=torch.randn(L, D, generator=g) * 0.01,
                                  transformer=Tower(), ln_final=torch.nn.Identity(),
                                  text_projection=torch.randn(D, 8, generator=g) * 0.1)
    sot, eot = _randint(g, 0, V - 1), _randint(g, 0, V - 1)
    clip = types.SimpleNamespace(model=model, device=torch.device("cpu"), selected_text_emb_ids=None,
                                 tokenizer=types.SimpleNamespace(encoder={"<|startoftext|>": sot, "<|endoftext|>": eot}))
    kw = (torch.randn(B, Kmax, D, generator=g) * 0.02).requires_grad_(True)
    if dynamic:
        num = torch.randint(0, Kmax + 1, (B,), generator=g)
        num[0] = Kmax                                # the edges: a full and (if B > 1) an empty keyword sequence
        if B > 1:
            num[1] = 0
        keyword_num = num
    else:
        keyword_num, num = Kmax, torch.full((B,), Kmax)
    out = co.ClipModel.encode_keywords(clip, kw, keyword_num)
    gout = torch.randn(out.shape, generator=g)
    (gk,) = torch.autograd.grad(out, [kw], grad_outputs=gout)
    kw2 = kw.detach().clone().requires_grad_(True)
    x, index = oracle.splice_keywords(kw2, keyword_num, emb.weight, model.positional_embedding, sot, eot)
    assert torch.equal(x, captured["x"])
    out2 = (x * 1.5)[torch.arange(B), index] @ model.text_projection
    assert rel_err(out2, out) < 1e-6
    (gk2,) = torch.autograd.grad(out2, [kw2], grad_outputs=gout)
    assert rel_err(gk2, gk) < 1e-6
    assert torch.equal(oracle.keypadding_mask(Kmax, num), du.get_keypadding_mask(Kmax, num))


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(N_CASES // 2))
def test_hybrid_compute_loss_live(refmods, seed):
    """C0: KWClip_GeneralTransformer.compute_loss (kwClip.py:999-1040) with random objective weights, either or both
    branches, with and without the CIF quantity loss."""
    ref.import_avssl()
    import avssl.model.kwClip as kc
    g = _gen(600 + seed)
    N, D = _randint(g, 2, 64), 8 * _randint(g, 1, 16)
    w_c = [0.0, 1.0, 1.5][_randint(g, 0, 2)]
    w_p = [1.0, 0.5][_randint(g, 0, 1)] if w_c == 0.0 else [0.0, 1.0, 0.5][_randint(g, 0, 2)]
    with_quantity = _coin(g)
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    ca = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + img, dim=-1).requires_grad_(True)
    pa = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + 2 * img, dim=-1).requires_grad_(True)
    ids = torch.randint(0, max(N // 3, 1), (N,), generator=g)
    fake = types.SimpleNamespace()
    fake.config = types.SimpleNamespace(model_settings=types.SimpleNamespace(cascaded_objective_weight=w_c,
                                                                             parallel_objective_weight=w_p))
    refmods.losses.MAX_EYE = 256
    fake.criterion = refmods.losses.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True)
    feats = {"id": ids, "image_feat": img, "cascaded_audio_feat": ca, "parallel_audio_feat": pa}
    q_w = 0.0
    if with_quantity:
        fake.quantity_loss_criteria = torch.nn.L1Loss()
        fake.quantity_loss_weight = q_w = float(torch.rand(1, generator=g))
        feats["cif_quantity_out"] = torch.rand(N, generator=g) * 10
        feats["cif_target_len"] = torch.randint(4, 12, (N,), generator=g).float()
    out = kc.KWClip_GeneralTransformer.compute_loss(fake, feats)
    params = [p for p, w in ((ca, w_c), (pa, w_p)) if w > 0] + [fake.criterion.temperature]
    grads = torch.autograd.grad(out["loss"], params)
    t = torch.tensor(math.log(1.0 / 0.07), requires_grad=True)
    ca2, pa2 = ca.detach().clone().requires_grad_(True), pa.detach().clone().requires_grad_(True)
    feats2 = dict(feats, cascaded_audio_feat=ca2, parallel_audio_feat=pa2)
    out2 = oracle.hybrid_loss(feats2, t.exp(), w_c, w_p, quantity_loss_weight=q_w)
    assert set(out2) == set(out), (sorted(out2), sorted(out))
    for key in out:
        assert rel_err(out2[key], out[key]) < 2e-5, key
    params2 = [p for p, w in ((ca2, w_c), (pa2, w_p)) if w > 0] + [t]
    for a, b in zip(torch.autograd.grad(out2["loss"], params2), grads):
        assert norm_err(a, b) < 1e-4
