"""GPU parity tests (run on the B200 box: pytest -m gpu).

Every test drives the CUDA path through the reference-shaped Python modules, which call the C ABI of
libscp_b200.so via ctypes, and compares with
  (1) the committed golden fixtures produced by the REFERENCE's own classes (tests/golden/*.npz),
  (2) the CPU oracle on seeded inputs (fp64 re-evaluation classifies arg-max ties),
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances (BASELINE.json north_star): VQ code indices bit-exact except exact ties; features / losses / gradients /
logged metrics within 1e-3 relative error (max |a-b| / max |b|, or ||a-b|| / ||b|| for gradients).
"""
import math
import types

import pytest
import torch

from conftest import golden_names, load_golden, norm_err, rel_err
from oracle import speechclip_oracle as oracle

pytestmark = pytest.mark.gpu

import os  # noqa: E402

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL = 1e-3


@pytest.fixture(scope="module")
def scp():
    import speechclip_plus_b200 as m
    m.load_library()
    return m


def _temp(spec: str) -> float:
    return float(spec.split("=")[1])


# =====================================================================================================================
# S1 weighted sum
# =====================================================================================================================
@pytest.mark.parametrize("name", golden_names("wsum_"))
def test_wsum_golden(scp, name):
    g = load_golden(name)
    L = g["layers_tbd"].shape[0]
    layer = scp.WeightedSumLayer(L, normalize_features=g["normalize"]).cuda()
    with torch.no_grad():
        layer.weights.copy_(g["weights"])
    # same memory layout as the reference hands over: (T,B,D) storage viewed as (B,T,D)
    layers = [t.cuda().transpose(0, 1).requires_grad_(True) for t in g["layers_tbd"]]
    y = layer(layers)
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    assert rel_err(y, g["y"]) < TOL
    grads = torch.autograd.grad(y, [layer.weights] + layers, grad_outputs=g["grad_y"].cuda())
    assert rel_err(grads[0], g["grad_weights"]) < TOL
    assert norm_err(torch.stack(grads[1:]), g["grad_layers"]) < TOL


@pytest.mark.parametrize("name", golden_names("s1tail_"))
def test_upstream_tail_golden(scp, name):
    """S1' through fuse_upstream_features vs what the reference's HuBERT wrapper forward returned."""
    from speechclip_plus_b200.module.speech_encoder_plus import fuse_upstream_features
    g = load_golden(name)
    L = g["layers_tbd"].shape[0]
    norm, ntype = g["normalize_hiddenstates"], g["normalize_type"]
    layer = scp.WeightedSumLayer(L, normalize_features=norm and ntype == "s3prl").cuda()  # speech_encoder_plus.py:472-476
    with torch.no_grad():
        layer.weights.copy_(g["weights"])
    trainable_upstream = ntype != "method2"  # layer gradients through method2 are not implemented (documented)
    layers = [t.cuda().transpose(0, 1).requires_grad_(trainable_upstream) for t in g["layers_tbd"]]
    y, feat_len = fuse_upstream_features(layers, layer, norm, ntype, g["wav_len"].tolist(), 320)
    assert torch.equal(feat_len.cpu(), g["feat_len"]) and feat_len.dtype == torch.int64 and feat_len.is_cuda
    assert rel_err(y, g["y"]) < TOL
    if trainable_upstream:
        grads = torch.autograd.grad(y, [layer.weights] + layers, grad_outputs=g["grad_y"].cuda())
        assert norm_err(torch.stack(grads[1:]), g["grad_layers"]) < TOL
    else:
        grads = torch.autograd.grad(y, [layer.weights], grad_outputs=g["grad_y"].cuda())
        with pytest.raises(scp.ScpError):  # method2 with a trainable upstream: loud, not silent
            xs = [t.cuda().transpose(0, 1).requires_grad_(True) for t in g["layers_tbd"]]
            y2, _ = fuse_upstream_features(xs, layer, norm, ntype)
            torch.autograd.grad(y2, xs[:1], grad_outputs=g["grad_y"].cuda())
    assert rel_err(grads[0], g["grad_weights"]) < TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("ntype", ["method1", "method2"])
@pytest.mark.parametrize("shape", [(13, 8, 249, 768), (25, 3, 61, 1024)])
def test_upstream_tail_vs_oracle(scp, shape, ntype, dtype):
    from speechclip_plus_b200.module.speech_encoder_plus import fuse_upstream_features
    L, B, T, D = shape
    gen = torch.Generator().manual_seed(L * 100 + T)
    storage = [(torch.randn(T, B, D, generator=gen) * (1 + 0.2 * l) + 0.05 * l).to(dtype) for l in range(L)]
    w = torch.randn(L, generator=gen) * 0.5
    gy = torch.randn(B, T, D, generator=gen)
    layer = scp.WeightedSumLayer(L).cuda()
    with torch.no_grad():
        layer.weights.copy_(w)
    wav_len = [320 * T - 37 * b for b in range(B)]
    y, feat_len = fuse_upstream_features([s.cuda().transpose(0, 1) for s in storage], layer, True, ntype, wav_len)
    (dw,) = torch.autograd.grad(y, [layer.weights], grad_outputs=gy.cuda())
    ref_layers = [s.double().transpose(0, 1) for s in storage]
    w_ref = w.double().requires_grad_(True)
    y_ref, len_ref = oracle.upstream_tail(ref_layers, w_ref, True, ntype, wav_len)
    (dw_ref,) = torch.autograd.grad(y_ref, [w_ref], grad_outputs=gy.double())
    assert torch.equal(feat_len.cpu(), len_ref)
    assert rel_err(y, y_ref) < TOL
    assert rel_err(dw, dw_ref) < TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("norm", [False, True])
@pytest.mark.parametrize("shape", [(13, 8, 249, 768), (25, 3, 61, 1024), (5, 2, 7, 64), (17, 2, 33, 256), (29, 2, 33, 512)])
def test_wsum_vs_oracle(scp, shape, norm, dtype):
    L, B, T, D = shape
    gen = torch.Generator().manual_seed(L * 1000 + T)
    storage = [(torch.randn(T, B, D, generator=gen) * (1 + 0.2 * l) + 0.05 * l).to(dtype) for l in range(L)]
    w = torch.randn(L, generator=gen) * 0.5
    gy = torch.randn(B, T, D, generator=gen)
    layer = scp.WeightedSumLayer(L, norm).cuda()
    with torch.no_grad():
        layer.weights.copy_(w)
    y = layer([s.cuda().transpose(0, 1) for s in storage])
    (dw,) = torch.autograd.grad(y, [layer.weights], grad_outputs=gy.cuda())
    ref_layers = [s.double().transpose(0, 1) for s in storage]
    y_ref = oracle.wsum_forward(ref_layers, w.double(), norm)
    dw_ref = oracle.wsum_grad_weights(ref_layers, w.double(), gy.double(), norm)
    assert rel_err(y, y_ref) < TOL
    assert rel_err(dw, dw_ref) < TOL


def test_wsum_contiguous_and_generic_inputs(scp):
    L, B, T, D = 4, 3, 5, 128
    gen = torch.Generator().manual_seed(1)
    xs = [torch.randn(B, T, D, generator=gen) for _ in range(L)]
    layer = scp.WeightedSumLayer(L).cuda()
    y = layer([x.cuda() for x in xs])  # plain contiguous (B,T,D)
    assert rel_err(y, oracle.wsum_forward(xs, torch.zeros(L))) < 1e-5
    y2 = layer([x.cuda()[:, :, :] if i else x.cuda().clone() for i, x in enumerate(xs)])
    assert torch.equal(y, y2)
    with pytest.raises(AssertionError):
        layer([x.cuda() for x in xs[:-1]])  # weighted_sum.py:36
    with pytest.raises(scp.ScpError):
        layer(xs)  # CPU tensors: there is no CPU path


def test_wsum_full_size_properties(scp):
    """BASELINE config 2/3 per-GPU size: L=13, B=256, T=249, D=768 (2.55 GB of layer features)."""
    L, B, T, D = 13, 256, 249, 768
    gen = torch.Generator(device="cuda").manual_seed(3)
    storage = [torch.randn(T, B, D, device="cuda", generator=gen) for _ in range(L)]
    layers = [s.transpose(0, 1) for s in storage]
    layer = scp.WeightedSumLayer(L).cuda()
    # (a) softmax weights sum to one: identical layers -> output equals the layer
    y = layer([layers[0]] * L)
    assert rel_err(y[::17, ::5], layers[0][::17, ::5]) < 1e-6
    # (b) a dominant weight selects its layer
    with torch.no_grad():
        layer.weights.zero_()
        layer.weights[7] = 60.0
    y = layer(layers)
    assert rel_err(y[::17, ::5], layers[7][::17, ::5]) < 1e-6
    # (c) linearity in the layer tensors and the closed-form weight gradient on a strided sample
    with torch.no_grad():
        layer.weights.copy_(torch.linspace(-1, 1, L))
    y = layer(layers)
    w = torch.softmax(layer.weights.detach(), 0)
    ref = sum(w[l] * layers[l][::31] for l in range(L))
    assert rel_err(y[::31], ref) < 1e-5
    gy = torch.randn(B, T, D, device="cuda", generator=gen)
    (dw,) = torch.autograd.grad(y, [layer.weights], grad_outputs=gy)
    d = torch.stack([(gy.double() * layers[l].double()).sum() for l in range(L)])
    dw_ref = w.double() * (d - (w.double() * d).sum())
    assert rel_err(dw, dw_ref) < TOL


# =====================================================================================================================
# N1 keyword batch-norm prologue
# =====================================================================================================================
def _build_kwbn(g, D, K):
    from speechclip_plus_b200.module.kw_bn import Kw_BatchNorm, Kw_BatchNorm_dynamic
    btype = g["batchnorm_type"]
    zeros, ones = torch.zeros(D), torch.ones(D)
    if btype == "dynamic":
        layer = Kw_BatchNorm_dynamic(kw_dim=D, init_bias=zeros, init_scale=ones)
    else:
        layer = Kw_BatchNorm(kw_num=K, kw_dim=D, batchnorm_type=btype, init_bias=zeros, init_scale=ones,
                             parallel=g["parallel"])
    bns = list(layer.bn_layers) if hasattr(layer, "bn_layers") else [layer.bn_layer]
    with torch.no_grad():
        for i, bn in enumerate(bns):
            bn.weight.copy_(g["weight"][i]); bn.bias.copy_(g["bias"][i])
            bn.running_mean.copy_(g["running_mean_in"][i]); bn.running_var.copy_(g["running_var_in"][i])
    layer.cuda().train(g["training"])
    return layer, bns


@pytest.mark.parametrize("name", golden_names("kwbn_"))
def test_kw_batchnorm_golden(scp, name):
    g = load_golden(name)
    B, K, D = g["x"].shape
    layer, bns = _build_kwbn(g, D, K)
    x = g["x"].cuda().requires_grad_(True)
    seq_lens = g["seq_lens"]
    y = layer(x, seq_lens.cuda()) if seq_lens.numel() else layer(x)
    assert y.shape == g["y"].shape
    assert rel_err(y, g["y"]) < TOL
    params = [p for bn in bns for p in (bn.weight, bn.bias)]
    grads = torch.autograd.grad(y, [x] + params, grad_outputs=g["grad_y"].cuda())
    assert norm_err(grads[0], g["grad_x"]) < TOL
    assert rel_err(torch.stack(list(grads[1::2])), g["grad_weight"]) < TOL
    assert rel_err(torch.stack(list(grads[2::2])), g["grad_bias"]) < TOL
    # running statistics and the batch counter follow torch.nn.BatchNorm1d
    assert rel_err(torch.stack([bn.running_mean for bn in bns]), g["running_mean_out"]) < TOL
    assert rel_err(torch.stack([bn.running_var for bn in bns]), g["running_var_out"]) < TOL
    assert torch.equal(torch.stack([bn.num_batches_tracked for bn in bns]).cpu(), g["num_batches_tracked"])


@pytest.mark.parametrize("cfg", [("eachKw", True, 256, 8, 512), ("same", False, 256, 8, 512), ("same", False, 64, 12, 768),
                                 ("eachKw", True, 3, 5, 64)])
def test_kw_batchnorm_vs_oracle(scp, cfg):
    from speechclip_plus_b200.module.kw_bn import Kw_BatchNorm
    btype, parallel, B, K, D = cfg
    gen = torch.Generator().manual_seed(B * 7 + K)
    x = (torch.randn(B, K, D, generator=gen) * 0.6 + 3.0 * torch.randn(1, 1, D, generator=gen))  # |mean| >> std
    init_bias, init_scale = torch.randn(D, generator=gen) * 0.01, torch.rand(D, generator=gen) * 0.02 + 0.01
    layer = Kw_BatchNorm(K, D, btype, init_bias, init_scale, parallel=parallel).cuda().train()
    bn = layer.bn_layer
    xd = x.cuda().requires_grad_(True)
    y = layer(xd)
    gy = torch.randn(B, K, D, generator=gen)
    gx, gw, gb = torch.autograd.grad(y, [xd, bn.weight, bn.bias], grad_outputs=gy.cuda())
    xr = x.double().requires_grad_(True)
    w_r = bn.weight.detach().double().cpu().requires_grad_(True)
    b_r = bn.bias.detach().double().cpu().requires_grad_(True)
    n_feat = w_r.numel()
    y_r, rm_r, rv_r = oracle.kw_batchnorm(xr, w_r, b_r, torch.zeros(n_feat, dtype=torch.float64),
                                          torch.ones(n_feat, dtype=torch.float64), btype, parallel, True)
    gx_r, gw_r, gb_r = torch.autograd.grad(y_r, [xr, w_r, b_r], grad_outputs=gy.double())
    assert rel_err(y, y_r) < TOL
    assert norm_err(gx, gx_r) < TOL and rel_err(gw, gw_r) < TOL and rel_err(gb, gb_r) < TOL
    assert rel_err(bn.running_mean, rm_r) < TOL and rel_err(bn.running_var, rv_r) < TOL
    # determinism: the slice partials are combined in a fixed order
    layer2 = Kw_BatchNorm(K, D, btype, init_bias, init_scale, parallel=parallel).cuda().train()
    assert torch.equal(layer2(x.cuda()), y)


def test_kw_batchnorm_state_dict_and_errors(scp):
    from speechclip_plus_b200.module.kw_bn import Kw_BatchNorm, Kw_BatchNorm_dynamic
    D, K = 64, 4
    par = Kw_BatchNorm(K, D, "eachKw", torch.zeros(D), torch.ones(D), parallel=True)
    assert sorted(par.state_dict()) == ["bn_layer.bias", "bn_layer.num_batches_tracked", "bn_layer.running_mean",
                                        "bn_layer.running_var", "bn_layer.weight"]
    assert par.bn_layer.weight.shape == (D * K,)                                 # kw_bn.py:46
    lay = Kw_BatchNorm(K, D, "eachKw", torch.zeros(D), torch.ones(D), parallel=False)
    assert "bn_layers.3.running_var" in lay.state_dict()                         # kw_bn.py:48-50
    dyn = Kw_BatchNorm_dynamic(D, torch.zeros(D), torch.ones(D), std_scale=2.0, learnable=False).cuda()
    assert not dyn.bn_layer.weight.requires_grad and float(dyn.bn_layer.weight[0]) == 2.0
    with pytest.raises(NotImplementedError):
        Kw_BatchNorm(K, D, "perRow", torch.zeros(D), torch.ones(D))
    with pytest.raises(AssertionError):
        par.cuda()(torch.zeros(2, K + 1, D, device="cuda"))                      # kw_bn.py:112
    with pytest.raises(ValueError):
        par.train()(torch.zeros(1, K, D, device="cuda"))                         # one value per channel in training
    with pytest.raises(scp.ScpError):
        dyn(torch.zeros(2, 3, D))                                                # CPU tensor: no fallback


# =====================================================================================================================
# N3 text-transformer input splice
# =====================================================================================================================
def _fake_clip(g, dtype=torch.float32):
    import types
    V, D = g["table"].shape
    emb = torch.nn.Embedding(V, D)
    emb.weight.data.copy_(g["table"])
    emb.weight.requires_grad_(False)
    emb = emb.cuda().to(dtype)
    captured = {}

    class Tower(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mix = torch.nn.Linear(D, D)
            self.mix.weight.data.copy_(g["mix_w"]); self.mix.bias.data.copy_(g["mix_b"])

        def forward(self, x):
            captured["x"] = x.permute(1, 0, 2)
            return torch.tanh(self.mix(x))

    model = types.SimpleNamespace(token_embedding=emb, positional_embedding=g["pos_emb"].cuda().to(dtype),
                                  transformer=Tower().cuda().to(dtype), ln_final=torch.nn.LayerNorm(D).cuda().to(dtype),
                                  text_projection=g["text_projection"].cuda().to(dtype))
    sot, eot = int(g["sot"]), int(g["eot"])
    clip = types.SimpleNamespace(model=model, device=torch.device("cuda"), selected_text_emb_ids=None,
                                 tokenizer=types.SimpleNamespace(encoder={"<|startoftext|>": sot, "<|endoftext|>": eot}))
    return clip, captured


@pytest.mark.parametrize("name", golden_names("splice_"))
def test_keyword_splice_golden(scp, name):
    from speechclip_plus_b200.module.clip_glue import encode_keywords, get_keypadding_mask
    g = load_golden(name)
    clip, captured = _fake_clip(g)
    if name.endswith("reduced"):  # token ids taken from the reduced-vocabulary attributes (clip_official.py:246-247)
        clip.selected_text_emb_ids = list(range(g["table"].shape[0]))
        clip.startOfTxt_reduced, clip.endOfTxt_reduced = int(g["sot"]), int(g["eot"])
        clip.tokenizer.encoder = {"<|startoftext|>": 0, "<|endoftext|>": 1}
    kw = g["keywords"].cuda().requires_grad_(True)
    num = g["keyword_num"]
    keyword_num = num.cuda() if num.dim() else int(num)
    out = encode_keywords(clip, kw, keyword_num)
    assert torch.equal(captured["x"].cpu(), g["x"])        # bit-exact: data movement + one fp32 add
    assert rel_err(out, g["out"]) < TOL
    (g_kw,) = torch.autograd.grad(out, [kw], grad_outputs=g["grad_out"].cuda())
    assert norm_err(g_kw, g["grad_keywords"]) < TOL
    lens = num if num.dim() else torch.full((kw.shape[0],), int(num))
    mask = get_keypadding_mask(kw.shape[1], lens.cuda())
    assert mask.dtype == torch.bool and torch.equal(mask.cpu(), g["keypadding_mask"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_keyword_splice_full_size(scp, dtype):
    """B=256, 77 x 512 CLIP text width (BASELINE config 3), dynamic keyword counts up to the CIF cap."""
    from speechclip_plus_b200.module.clip_glue import splice_keywords
    B, Kmax, D, V, L = 256, 75, 512, 49408, 77
    gen = torch.Generator(device="cuda").manual_seed(5)
    table = (torch.randn(V, D, device="cuda", generator=gen) * 0.02).to(dtype)
    pos = (torch.randn(L, D, device="cuda", generator=gen) * 0.01).to(dtype)
    kw = (torch.randn(B, Kmax, D, device="cuda", generator=gen) * 0.02).requires_grad_(True)
    num = torch.randint(0, Kmax + 1, (B,), device="cuda", generator=gen)
    x, idx = splice_keywords(kw, num, table, pos, V - 2, V - 1)
    x_ref, idx_ref = oracle.splice_keywords(kw.detach().cpu(), num.cpu(), table.cpu(), pos.cpu(), V - 2, V - 1)
    assert x.dtype == dtype and torch.equal(idx.cpu(), idx_ref)
    assert torch.equal(x.cpu(), x_ref)
    gx = torch.randn(B, L, D, device="cuda", generator=gen).to(dtype)
    (g_kw,) = torch.autograd.grad(x, [kw], grad_outputs=gx)
    valid = (torch.arange(Kmax, device="cuda")[None, :] < num[:, None])[..., None]
    assert torch.equal(g_kw, torch.where(valid, gx[:, 1:1 + Kmax].float(), torch.zeros((), device="cuda")))
    with pytest.raises(RuntimeError):
        splice_keywords(kw, 8, table, pos, V - 2, V - 1)              # int keyword_num must equal keywords.shape[1]
    with pytest.raises(scp.ScpError):
        splice_keywords(kw, num, table.clone().requires_grad_(True), pos, V - 2, V - 1)


# =====================================================================================================================
# N4 CIF down-sampler
# =====================================================================================================================
def _scaled_alpha(raw, target):
    if target is None:
        return raw
    return raw * ((1.0 * target.type_as(raw) + 1e-5) / raw.sum(1)).unsqueeze(1)   # cif.py:126-129


@pytest.mark.parametrize("name", [n for n in golden_names("cif_") if n != "cif_forward_train"])
def test_cif_integrate_and_fire_golden(scp, name):
    from speechclip_plus_b200.module.cif import integrate_and_fire
    g = load_golden(name)
    x = g["x"].cuda().requires_grad_(True)
    raw = g["alpha_raw"].cuda().requires_grad_(True)
    target = g["target_len"].cuda() if g["target_len"].numel() else None
    res = integrate_and_fire(x, _scaled_alpha(raw, target), 1.0, target, g["apply_tail_handling"])
    assert torch.equal(res["dsample_feats_length"].cpu(), g["feat_len"])
    assert torch.equal(res["fired_marks"].cpu(), g["fired_marks"])
    assert torch.equal(res["dsample_feats_pad_mask"].cpu(), g["pad_mask"])
    assert res["dsample_feats"].shape == g["feats"].shape and rel_err(res["dsample_feats"], g["feats"]) < TOL
    if g["mode"] != "infer":
        gx, ga = torch.autograd.grad(res["dsample_feats"], [x, raw], grad_outputs=g["grad_feats"].cuda())
        assert norm_err(gx, g["grad_x"]) < TOL and norm_err(ga, g["grad_alpha_raw"]) < TOL


def test_cif_module_forward_golden(scp):
    """Full CIF.forward (conv weight generator + clip / mask + scaling + integrate) with the reference's parameters."""
    from speechclip_plus_b200.module.cif import CIF
    g = load_golden("cif_forward_train")
    C = g["x"].shape[-1]
    layer = CIF(cif_threshold=1.0, cif_output_dim=C, encoder_embed_dim=C, conv_cif_width=3, scaling_step=100)
    assert sorted(n.replace(".", "_") for n, _ in layer.named_parameters()) == \
        sorted(k[len("param_"):] for k in g if k.startswith("param_"))          # same parameter names as the reference
    with torch.no_grad():
        for n, p in layer.named_parameters():
            p.copy_(g["param_" + n.replace(".", "_")])
    layer = layer.cuda().eval()
    x = g["x"].cuda().requires_grad_(True)
    res = layer({"audio_feat": x, "audio_feat_pad_mask": g["pad_mask_in"].cuda(), "global_step": 0},
                g["target_len"].cuda())
    assert torch.equal(res["dsample_feats_length"].cpu(), g["feat_len"])
    assert torch.equal(res["original_length"].cpu(), g["original_length"])
    for key, ref_key in (("dsample_feats", "feats"), ("quantity_out", "quantity_out"), ("orig_alpha", "orig_alpha"),
                         ("alpha", "alpha")):
        assert rel_err(res[key], g[ref_key]) < TOL, key
    params = list(layer.parameters())
    grads = torch.autograd.grad(res["dsample_feats"].mul(g["grad_feats"].cuda()).sum() + res["quantity_out"].sum(),
                                [x] + params)
    assert norm_err(grads[0], g["grad_x"]) < TOL
    for (n, _), gr in zip(layer.named_parameters(), grads[1:]):
        assert norm_err(gr, g["grad_" + n.replace(".", "_")]) < 2e-3, n
    assert layer.apply_scaling                                           # global_step 0 < scaling_step (cif.py:102-104)
    layer({"audio_feat": x.detach(), "audio_feat_pad_mask": g["pad_mask_in"].cuda(), "global_step": 100}, g["target_len"].cuda())
    assert not layer.apply_scaling


@pytest.mark.parametrize("mode", ["train", "infer"])
def test_cif_full_size_vs_oracle(scp, mode):
    """B=64 utterances of 249 HuBERT frames x 768 (the "+" branches' input), ~12 keywords each."""
    from speechclip_plus_b200.module.cif import integrate_and_fire
    B, S, C = 64, 249, 768
    gen = torch.Generator().manual_seed(11 if mode == "train" else 12)
    x = torch.randn(B, S, C, generator=gen)
    raw = torch.rand(B, S, generator=gen) * 0.1
    lens = torch.randint(120, S + 1, (B,), generator=gen)
    raw = raw.masked_fill(torch.arange(S)[None, :] >= lens[:, None], 0.0)
    target = (lens / 20).round().long() if mode == "train" else None       # kw_branches.py:684
    xd, rd = x.cuda().requires_grad_(True), raw.cuda().requires_grad_(True)
    res = integrate_and_fire(xd, _scaled_alpha(rd, target.cuda() if target is not None else None), 1.0,
                             target.cuda() if target is not None else None)
    xr, rr = x.double().requires_grad_(True), raw.double().requires_grad_(True)
    feats, feat_len, fired = oracle.cif_integrate_and_fire(xr, _scaled_alpha(rr, target), 1.0, target)
    assert torch.equal(res["dsample_feats_length"].cpu(), feat_len)
    assert torch.equal(res["fired_marks"].cpu(), fired)
    assert rel_err(res["dsample_feats"], feats) < TOL
    gy = torch.randn(feats.shape, generator=gen)
    gx, ga = torch.autograd.grad(res["dsample_feats"], [xd, rd], grad_outputs=gy.cuda())
    gx_r, ga_r = torch.autograd.grad(feats, [xr, rr], grad_outputs=gy.double())
    assert norm_err(gx, gx_r) < TOL and norm_err(ga, ga_r) < TOL
    # conservation: without scaling every unit of alpha lands in exactly one output row (incl. the tail)
    if mode == "infer":
        ones = torch.ones(B, S, 4, device="cuda")
        r1 = integrate_and_fire(ones, rd.detach(), 1.0, None, apply_tail_handling=False)
        assert (r1["dsample_feats"][..., 0].sum(1) <= rd.detach().sum(1) + 1e-3).all()


# =====================================================================================================================
# S2 vector quantiser
# =====================================================================================================================
def _make_vq(scp, spec, training, hard=True):
    vq = scp.SimpleVectorQuantizer(spec, hard=hard).cuda()
    vq.train(training)
    return vq


class _BranchStub:
    """What fused_vq_audio_features needs of a GeneralBranch: identity projection, the frozen table and the reference's
    dense cosine scores (kw_branches.py:158-179, restated with one broadcast instead of the per-keyword loop)."""

    def __init__(self, vq, table):
        emb = torch.nn.Embedding.from_pretrained(table.clone(), freeze=True)
        self.vector_quantizer, self.clip = vq, types.SimpleNamespace(model=types.SimpleNamespace(token_embedding=emb))

    def project_feats_to_CLIPspace(self, feat):
        return feat

    def get_keyword_cosine_score(self, feat):
        table = self.clip.model.token_embedding.weight
        return torch.nn.functional.cosine_similarity(feat[:, :, None, :], table[None, None], dim=-1)


@pytest.mark.parametrize("name", golden_names("vq_"))
def test_vq_fused_golden(scp, name):
    g = load_golden(name)
    hard = bool(g.get("hard", True))
    vq = _make_vq(scp, g["temp_spec"], g["training"], hard)
    kw = g["keywords_in"].cuda().requires_grad_(True)
    if hard:
        res, out = vq.quantize_keywords(kw, g["table"].cuda())
    else:  # hard=False has no fused kernel: the drop-in body of vq_audio_features routes it through the dense quantiser
        with pytest.raises(NotImplementedError):
            vq.quantize_keywords(kw, g["table"].cuda())
        res, out = scp.fused_vq_audio_features(_BranchStub(vq, g["table"].cuda()), kw)
        assert rel_err(res["subword_prob"], g["subword_prob"]) < TOL
    # code indices: bit-exact (this includes the fixture with exactly tied duplicate table rows and a masked best match)
    assert torch.equal(res["targets"].cpu(), g["targets"])
    assert res["num_vars"] == int(g["num_vars"])
    assert rel_err(out, g["keywords_out"]) < TOL
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t", "diversity_loss"):
        assert rel_err(res[key], g[key]) < TOL, key
    assert math.isclose(res["temp"], float(g["temp"]), rel_tol=1e-6)
    if g["training"]:
        params = [kw] + ([vq.curr_temp] if g["temp_spec"].startswith("learnable") else [])
        grads = torch.autograd.grad(out, params, grad_outputs=g["grad_keywords_out"].cuda())
        assert norm_err(grads[0], g["grad_keywords_in"]) < TOL
        assert rel_err(grads[0], g["grad_keywords_in"]) < 2 * TOL
        if len(grads) > 1:
            # the reference hands NaN to a learnable temperature (d(-inf/tau)); we return the finite closed form
            _, g_tau = oracle.vq_keyword_grad(g["keywords_in"].double(), g["table"].double(),
                                              torch.tensor(_temp(g["temp_spec"]), dtype=torch.float64),
                                              g["grad_keywords_out"].double())
            assert torch.isfinite(grads[1]).all()
            assert rel_err(grads[1].reshape(()), g_tau) < 5 * TOL
    else:
        assert not out.requires_grad  # eval: one-hot lookup, no gradient path


@pytest.mark.parametrize("name", golden_names("vq_"))
def test_vq_dense_golden(scp, name):
    """The reference signature SimpleVectorQuantizer.forward(x) on the dense cosine scores."""
    g = load_golden(name)
    hard = bool(g.get("hard", True))
    vq = _make_vq(scp, g["temp_spec"], g["training"], hard)
    x = g["cos"].cuda().clone().requires_grad_(g["training"])
    xin = x.clone() if g["training"] else x  # a leaf that requires grad cannot be modified in place
    res = vq(xin)
    assert torch.equal(res["targets"].cpu(), g["targets"])
    assert torch.isinf(xin.detach()[..., [0, 2, 3]]).all()  # masked in place like my_vector_quantizer.py:78-79
    assert rel_err(res["subword_prob"], g["subword_prob"]) < TOL
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t", "diversity_loss"):
        assert rel_err(res[key], g[key]) < TOL, key
    kw_out = res["subword_prob"] @ g["table"].cuda()  # kw_branches.py:195
    assert rel_err(kw_out, g["keywords_out"]) < TOL
    if g["training"]:
        (gx,) = torch.autograd.grad(kw_out, [x], grad_outputs=g["grad_keywords_out"].cuda())
        # reference gradient w.r.t. the cosine scores via the oracle's autograd
        xr = g["cos"].clone().requires_grad_(True)
        vr = oracle.vq_forward(xr, torch.tensor([_temp(g["temp_spec"])]), training=True, hard=hard)
        (gr,) = torch.autograd.grad(vr["subword_prob"] @ g["table"], [xr], grad_outputs=g["grad_keywords_out"])
        gr = torch.nan_to_num(gr, nan=0.0)
        assert norm_err(gx, gr) < TOL


def _tie_tolerant_index_check(idx, kw, table, prob_msk=(0, 2, 3)):
    """Bit-exact except exact ties: a mismatch is tolerated only if the fp64 cosine gap is below fp32 summation noise."""
    cos = oracle.cosine_scores(kw.double(), table.double())
    cos[..., list(prob_msk)] = float("-inf")
    flat = cos.reshape(-1, cos.shape[-1])
    ref = flat.argmax(-1)
    idx = idx.reshape(-1).cpu()
    bad = (idx != ref).nonzero().flatten().tolist()
    for m in bad:
        gap = (flat[m, ref[m]] - flat[m, idx[m]]).item()
        assert gap < 3e-7, f"row {m}: picked {idx[m].item()} instead of {ref[m].item()}, cosine gap {gap:.3e}"
    return len(bad)


@pytest.mark.parametrize("tau", [0.1, 0.07])  # 0.1: the (e^c)^10 fast path of sweep 1; 0.07: the online-max path
@pytest.mark.parametrize("shape", [(32, 8, 8112, 512), (16, 12, 19787, 768), (7, 5, 1000, 64), (1, 1, 300, 64),
                                   (5, 3, 2000, 1024), (40, 8, 4096, 256), (33, 1, 600, 128), (3, 7, 12000, 384)])
def test_vq_fused_vs_oracle(scp, shape, tau):
    B, K, V, D = shape
    gen = torch.Generator().manual_seed(V + B)
    table = torch.randn(V, D, generator=gen) * 0.02 + 0.003 * torch.randn(1, D, generator=gen)
    kw = torch.randn(B, K, D, generator=gen) * table.std(0) + table.mean(0)
    gout = torch.randn(B, K, D, generator=gen)
    vq = _make_vq(scp, f"fixed={tau}", True)
    kwd = kw.cuda().requires_grad_(True)
    res, out = vq.quantize_keywords(kwd, table.cuda())
    n_ties = _tie_tolerant_index_check(res["targets"], kw, table)
    assert n_ties == 0  # random data has no 3e-7 near-ties; the exact re-scoring must reproduce the fp64 arg-max
    ref, out_ref = oracle.vq_audio_features(kw.double(), table.double(), torch.tensor([tau], dtype=torch.float64))
    assert rel_err(out, out_ref) < TOL
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t", "diversity_loss"):
        assert rel_err(res[key], ref[key]) < TOL, key
    assert rel_err(res["avg_probs"], ref["avg_probs"]) < TOL
    (gk,) = torch.autograd.grad(out, [kwd], grad_outputs=gout.cuda())
    g_ref, _ = oracle.vq_keyword_grad(kw.double(), table.double(), torch.tensor(tau, dtype=torch.float64), gout.double())
    assert norm_err(gk, g_ref) < TOL


@pytest.mark.parametrize("eps", [1e-2, 1e-4, 1e-6])
def test_vq_argmax_among_near_ties(scp, eps):
    """Stress of the exact arg-max: every table row is one of 16 directions plus a perturbation of relative size eps, so
    ~V/16 columns fall inside the fp16 rescue margin of each row's maximum (eps <= 1e-4: inside the fp32 margin as well)
    and all three re-scoring levels decide the winner.  The code must equal the fp64 arg-max of the fp32 inputs."""
    B, K, V, D = 16, 4, 2048 + 40, 256
    gen = torch.Generator().manual_seed(int(1 / eps))
    base = torch.nn.functional.normalize(torch.randn(16, D, generator=gen), dim=-1)
    table = base[torch.randint(0, 16, (V,), generator=gen)] * (1.0 + 0.5 * torch.rand(V, 1, generator=gen))
    table = table + eps * torch.randn(V, D, generator=gen) / D ** 0.5
    kw = base[torch.randint(0, 16, (B * K,), generator=gen)] + 0.05 * torch.randn(B * K, D, generator=gen)
    kw = kw.view(B, K, D)
    vq = _make_vq(scp, "fixed=0.1", False)
    res, out = vq.quantize_keywords(kw.cuda(), table.cuda())
    n_ties = _tie_tolerant_index_check(res["targets"], kw, table)
    assert n_ties <= 1            # only a gap below fp32 summation noise may differ (asserted inside the helper)
    assert torch.equal(out.cpu().view(-1, D), table[res["targets"].view(-1).cpu()])


def test_vq_edge_cases(scp):
    V, D = 520, 64
    gen = torch.Generator().manual_seed(5)
    table = torch.randn(V, D, generator=gen) * 0.02
    table[100] = 0.0  # a zero row: cosine 0 for every keyword (F.cosine_similarity eps clamp)
    kw = torch.randn(2, 3, D, generator=gen)
    kw[0, 1] = 0.0     # a zero keyword: all cosines 0 -> first unmasked column wins, like torch.max
    kw[1, 0] = table[3] * 5   # best raw match masked -> must pick something else
    kw[1, 2] = table[519] * 2  # last column
    vq = _make_vq(scp, "fixed=0.1", False)
    res, out = vq.quantize_keywords(kw.cuda(), table.cuda())
    ref, out_ref = oracle.vq_audio_features(kw.double(), table.double(), torch.tensor([0.1], dtype=torch.float64),
                                            training=False)
    assert torch.equal(res["targets"].cpu(), ref["targets"])
    assert res["targets"][0, 1, 0].item() == 1 and res["targets"][1, 2, 0].item() == 519
    assert res["targets"][1, 0, 0].item() not in (0, 2, 3)
    assert rel_err(out, out_ref) < 1e-6
    # custom mask list and empty mask
    res2, _ = vq.quantize_keywords(kw.cuda(), table.cuda(), prob_msk=())
    ref2, _ = oracle.vq_audio_features(kw.double(), table.double(), torch.tensor([0.1], dtype=torch.float64),
                                       training=False, prob_msk=())
    assert torch.equal(res2["targets"].cpu(), ref2["targets"])
    assert res2["targets"][1, 0, 0].item() == 3
    with pytest.raises(IndexError):
        vq.quantize_keywords(kw.cuda(), table.cuda(), prob_msk=(0, 9999))


@pytest.mark.parametrize("cfg", [(256, 8, 49408, 512), (128, 8, 49408, 768), (128, 12, 49408, 512)],
                         ids=["c3_base_M2048_D512", "c5_large_M1024_D768", "c4_dynamicK_M1536_D512"])
def test_vq_full_size_properties(scp, cfg):
    """BASELINE configs 3 / 5 / 4 per GPU: M = B*K keyword rows against the full 49408-row CLIP table (D = 512: resident-X
    kernels; D = 768: the streaming kernels)."""
    B, K, V, D = cfg
    gen = torch.Generator(device="cuda").manual_seed(11)
    table = torch.randn(V, D, device="cuda", generator=gen) * 0.02
    kw = torch.randn(B, K, D, device="cuda", generator=gen) * 0.02
    # plant exact answers in a quarter of the rows: kw = scale * table[r]  ->  idx == r (cosine exactly maximal)
    planted = torch.randint(4, V, (B * K // 4,), device="cuda", generator=gen)
    rows = torch.arange(0, B * K, 4, device="cuda")
    kw.view(-1, D)[rows] = table[planted] * 3.0
    kw.requires_grad_(True)
    vq = _make_vq(scp, "fixed=0.1", True)
    res, out = vq.quantize_keywords(kw, table)
    idx = res["targets"].view(-1)
    assert torch.equal(idx[rows], planted)
    assert not bool(((idx == 0) | (idx == 2) | (idx == 3)).any())          # masked codes never win
    assert torch.equal(out.detach().view(-1, D), table[idx])                # lookup is an exact gather
    assert abs(res["avg_probs"].sum().item() - 1.0) < 1e-4                  # mean of softmax rows sums to one
    assert res["code_hist"].sum().item() == B * K
    assert 1.0 <= res["code_perplexity"].item() <= B * K + 1e-3
    assert res["prob_perplexity"].item() <= V
    # rows are independent: a permutation of the rows permutes the codes
    perm = torch.randperm(B * K, device="cuda", generator=gen)
    res_p, _ = vq.quantize_keywords(kw.detach().view(-1, D)[perm].view(B, K, D), table)
    assert torch.equal(res_p["targets"].view(-1), idx[perm])
    # exact arg-max on a random sample of rows (fp64 on the GPU)
    sample = torch.randint(0, B * K, (64,), device="cuda", generator=gen)
    ks = kw.detach().view(-1, D)[sample].double()
    cos = (ks / ks.norm(dim=1, keepdim=True)) @ (table.double() / table.double().norm(dim=1, keepdim=True)).t()
    cos[:, [0, 2, 3]] = float("-inf")
    assert torch.equal(cos.argmax(-1), idx[sample])
    # the gradient through the normalisation is orthogonal to the keyword
    gout = torch.randn(B, K, D, device="cuda", generator=gen)
    (gk,) = torch.autograd.grad(out, [kw], grad_outputs=gout)
    dots = (gk * kw.detach()).sum(-1).abs().max().item()
    assert dots < 1e-3 * (gk.norm(dim=-1) * kw.detach().norm(dim=-1)).max().item()
    assert torch.isfinite(gk).all()


@pytest.mark.parametrize("cfg", [(256, 8, 49408, 512), (128, 8, 49408, 768), (128, 12, 49408, 512)],
                         ids=["c3_base_M2048_D512", "c5_large_M1024_D768", "c4_dynamicK_M1536_D512"])
def test_vq_full_size_vs_fp64_oracle(scp, cfg):
    """BASELINE configs 3 / 5 / 4 at their per-GPU sizes against the oracle itself (not properties): the oracle's
    restatement of my_vector_quantizer.py:64-165 + kw_branches.py:158-197 is evaluated in fp64 ON THE GPU (the (M,V)
    fp64 matrices of the reference algorithm need ~6 GB, which no CPU test box should be asked for): all indices
    (tie-tolerant), code / prob perplexity, ent_per_t, diversity_loss, avg_probs, the looked-up keywords and the
    keyword gradient.  The fp16 e^c scratch, the closed-form prob_perplexity renormalisation and the depth of the
    split-K / pipeline partial sums all change with M and V: this is where they are verified."""
    B, K, V, D = cfg
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(V + B * K + D)
    table = torch.randn(V, D, device=dev, generator=gen) * 0.02 + 0.003 * torch.randn(1, D, device=dev, generator=gen)
    kw = torch.randn(B, K, D, device=dev, generator=gen) * table.std(0) + table.mean(0)
    flat = kw.view(-1, D)
    n_peak = flat[::7].shape[0]    # a seventh of the rows sit on a table row: sharply peaked softmax_tau rows
    flat[::7] = table[torch.randint(4, V, (n_peak,), device=dev, generator=gen)] * 1.7 + 0.002 * torch.randn(
        n_peak, D, device=dev, generator=gen)
    gout = torch.randn(B, K, D, device=dev, generator=gen)
    tau = 0.1
    vq = _make_vq(scp, f"fixed={tau}", True)
    kwd = kw.clone().requires_grad_(True)
    res, out = vq.quantize_keywords(kwd, table)
    (gk,) = torch.autograd.grad(out, [kwd], grad_outputs=gout)
    t64 = torch.tensor([tau], dtype=torch.float64, device=dev)
    ref, out_ref = oracle.vq_audio_features(kw.double(), table.double(), t64, training=True)
    # indices: bit-exact except ties below fp32 summation noise (judged on the oracle's own fp64 scores)
    idx = res["targets"].view(-1)
    ridx = ref["targets"].view(-1)
    bad = (idx != ridx).nonzero().flatten()
    scores = ref["masked_scores"].view(-1, V)
    for m in bad.tolist():
        gap = (scores[m, ridx[m]] - scores[m, idx[m]]).item()
        assert gap < 3e-7, f"row {m}: picked {idx[m].item()} instead of {ridx[m].item()}, cosine gap {gap:.3e}"
    assert bad.numel() <= 2
    assert rel_err(out, out_ref) < TOL if bad.numel() == 0 else True
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t", "diversity_loss"):
        assert rel_err(res[key], ref[key]) < TOL, key
    assert rel_err(res["avg_probs"], ref["avg_probs"]) < TOL
    assert norm_err(res["avg_probs"], ref["avg_probs"]) < TOL
    del ref, out_ref, scores
    g_ref, _ = oracle.vq_keyword_grad(kw.double(), table.double(), t64.reshape(()), gout.double())
    assert norm_err(gk, g_ref) < TOL
    # row-wise as well: a peaked row's gradient is orders of magnitude smaller than a diffuse row's and would hide in
    # the global norm (floor: 1e-3 of the largest row norm)
    gr = g_ref.view(-1, D)
    row = (gk.view(-1, D).double() - gr).norm(dim=1) / (gr.norm(dim=1) + 1e-3 * gr.norm(dim=1).max())
    assert row.max().item() < 5 * TOL, row.max().item()


@pytest.mark.parametrize("cfg", [(32, 8, 8112, 512, 0.1), (24, 12, 19787, 256, 0.5), (256, 8, 49408, 512, 0.1), (20, 8, 8112, 768, 0.1)],
                         ids=["flickr_vocab", "generic_tau_D256", "bench_shape", "D768_streamed_ghat"])
def test_vq_saved_numerators_vs_recompute(scp, cfg):
    """The two forms of the straight-through backward (kw_branches.py:181-197 + my_vector_quantizer.py:130-136) must agree:
    `save_probs=True` (scp_vq_fwd_save / scp_vq_bwd_saved: the forward keeps P'' = exp((c-1)/tau + 10) as fp16, the column
    sums and the arg-max filter read it, the backward forms only g . E^T) against `save_probs=False` (the backward recomputes
    k . E^T and the soft-max).  Same indices, metrics within 1e-4 of each other, keyword gradients within 5e-4 of each other
    and within the 1e-3 tolerance of the fp64 oracle.  D = 768 has no resident tile: sweep 1 and sweep T stream the
    keyword / gradient tile through the ring instead."""
    from speechclip_plus_b200.module.vector_quantizers import _FusedVQFn
    B, K, V, D, tau = cfg
    gen = torch.Generator().manual_seed(B * 7 + V)
    table = (torch.randn(V, D, generator=gen) * 0.02 + 0.003 * torch.randn(1, D, generator=gen)).cuda()
    kw = (torch.randn(B, K, D, generator=gen).cuda() * table.std(0) + table.mean(0))
    pick = torch.randint(0, V, (B,), generator=gen).cuda()
    kw[:, 0] = table[pick] * 1.5 + 0.002 * torch.randn(B, D, generator=gen).cuda()   # rows with a cosine close to 1
    kw[:, 1] = -table[pick.flip(0)] * 0.7     # ... and close to -1: numerators at the bottom of the fp16 range (e^-10 at tau = 0.1)
    gout = torch.randn(B, K, D, generator=gen).cuda()
    vq = _make_vq(scp, f"fixed={tau}", True)
    cache = vq._table_cache.get(table)
    outs = []
    for save in (True, False):
        kwd = kw.clone().requires_grad_(True)
        out, idx, metrics, row_stats, code_hist, avg_probs = _FusedVQFn.apply(kwd, vq.curr_temp, cache, (0, 2, 3), True, True, save)
        (g,) = torch.autograd.grad(out, [kwd], grad_outputs=gout)
        outs.append((idx, metrics, avg_probs[:V], g))
    (i1, m1, a1, g1), (i0, m0, a0, g0) = outs
    assert torch.equal(i1, i0)
    assert rel_err(m1, m0) < 1e-4 and rel_err(a1, a0) < 2e-4
    assert norm_err(g1, g0) < 5e-4, norm_err(g1, g0)
    if V <= 20000:
        g_ref, _ = oracle.vq_keyword_grad(kw.double().cpu(), table.double().cpu(), torch.tensor(tau, dtype=torch.float64),
                                          gout.double().cpu())
        assert norm_err(g1, g_ref) < TOL and norm_err(g0, g_ref) < TOL


@pytest.mark.parametrize("mode", ["1", "2"], ids=["fused_ring_in_L2", "two_launches"])
def test_vq_backward_pipeline_opt_in(mode):
    """The producer/consumer form of the VQ backward (csrc/scp_vq_pipe.cuh, SCP_VQ_BWD_PIPE=1/2: MN-major tcgen05 operands,
    ring hand-off through release/acquire counters) against the fp64 oracle on the GPU, incl. the learnable-temperature
    gradient.  The mode is read once per process, hence the subprocess; tools/vq_bwd_check.py exits non-zero when the keyword
    gradient of any shape is off by more than 1e-3 (kw_branches.py:181-197 + my_vector_quantizer.py:130-136)."""
    import subprocess
    import sys
    env = dict(os.environ, SCP_VQ_BWD_PIPE=mode)
    for extra in ([], ["--learnable"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT_DIR, "tools", "vq_bwd_check.py"), "--shapes", "small", *extra],
                           env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert f'"mode": "{mode}"' in r.stdout


# =====================================================================================================================
# S3 masked InfoNCE, N0 normalise + pack, C0 compute_loss
# =====================================================================================================================
def _crit_from_golden(scp, g):
    return scp.MaskedContrastiveLoss(temperature=float(g["temperature"]), temperature_trainable=g["trainable"],
                                     margin=float(g["margin"]), dcl=g["dcl"], a2b=g["a2b"], b2a=g["b2a"]).cuda()


@pytest.mark.parametrize("name", golden_names("nce_"))
def test_nce_golden(scp, name):
    g = load_golden(name)
    crit = _crit_from_golden(scp, g)
    a = g["feat_a"].cuda().requires_grad_(True)
    b = g["feat_b"].cuda().requires_grad_(True)
    ids = g["ids"].cuda() if g["has_ids"] else None
    loss = crit(a, b, ids)
    assert loss.dim() == 0
    assert rel_err(loss, g["loss"]) < TOL
    assert math.isclose(crit.current_temperature, float(g["current_temperature"]), rel_tol=1e-5)
    params = [a, b] + ([crit.temperature] if g["trainable"] else [])
    grads = torch.autograd.grad(loss, params)
    assert norm_err(grads[0], g["grad_a"]) < TOL
    assert norm_err(grads[1], g["grad_b"]) < TOL
    if g["trainable"]:
        assert rel_err(grads[2], g["grad_temperature"]) < TOL


@pytest.mark.parametrize("N,D", [(256, 512), (1024, 512), (512, 768), (2048, 512), (33, 64)])
def test_nce_vs_oracle(scp, N, D):
    gen = torch.Generator().manual_seed(N + D)
    a = torch.nn.functional.normalize(torch.randn(N, D, generator=gen), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(N, D, generator=gen) + 0.5 * a, dim=-1)
    ids = torch.randint(0, max(N // 5, 3), (N,), generator=gen)
    crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True).cuda()
    ad, bd = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    loss = crit(ad, bd, ids.cuda())
    ga, gb, gt = torch.autograd.grad(loss, [ad, bd, crit.temperature])
    scale = 1 / 0.07
    l_ref = oracle.nce_forward(a.double(), b.double(), ids, scale)
    da, db, dl = oracle.nce_grads(a.double(), b.double(), ids, scale)
    assert rel_err(loss, l_ref) < TOL
    assert norm_err(ga, da) < TOL and norm_err(gb, db) < TOL
    assert rel_err(gt, dl) < TOL


@pytest.mark.parametrize("N,D,world", [(256, 512, 2), (1024, 512, 8), (2048, 512, 8), (512, 768, 4), (66, 64, 2)])
def test_nce_sharded_forward_matches_full(scp, N, D, world):
    """SURVEY section 8(e) option B on ONE device: every rank's shard call (scp_nce_fwd_local on rows [r n, (r+1) n)) is
    issued in turn, the (3, n) statistics are concatenated as the all-gather would deliver them, and
    scp_nce_loss_from_stats must reproduce the loss / lse_row / lse_col of the full forward (losses.py:224-243)."""
    from speechclip_plus_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(N + D)
    a = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=gen), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=gen), dim=-1)
    ids = torch.randint(0, max(N // 5, 2), (N,), device=dev, generator=gen)
    ls = torch.tensor([math.log(1 / 0.07)], device=dev)
    n = N // world
    stream = _lib.stream_ptr(dev)
    ws_b = lib.scp_nce_workspace_bytes(N, D)
    ws = torch.empty(ws_b, dtype=torch.uint8, device=dev)
    loss_f = torch.empty(1, device=dev); lr_f = torch.empty(N, device=dev); lc_f = torch.empty(N, device=dev)
    _lib.check(lib.scp_nce_fwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ids), N, D, _lib.ptr(ls), 0.0, 0.0, 0, 1, 1, 0,
                               _lib.ptr(loss_f), _lib.ptr(lr_f), _lib.ptr(lc_f), _lib.ptr(ws), ws_b, stream), "fwd")
    stats_all = torch.empty((world, 3, n), device=dev)
    for r in range(world):
        _lib.check(lib.scp_nce_fwd_local(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ids), N, D, _lib.ptr(ls), 0.0, 0.0, 0,
                                         r * n, (r + 1) * n, 0, _lib.ptr(stats_all[r]), _lib.ptr(ws), ws_b, stream),
                   "fwd_local")
    loss_s = torch.empty(1, device=dev); lr_s = torch.empty(N, device=dev); lc_s = torch.empty(N, device=dev)
    _lib.check(lib.scp_nce_loss_from_stats(_lib.ptr(stats_all), world, n, _lib.ptr(ls), 0.0, 0.0, 1, 1,
                                           _lib.ptr(loss_s), _lib.ptr(lr_s), _lib.ptr(lc_s), stream), "loss_from_stats")
    torch.cuda.synchronize()
    ref = oracle.nce_forward(a.double().cpu(), b.double().cpu(), ids.cpu(), 1 / 0.07)
    assert rel_err(loss_s, ref.reshape(1)) < 1e-5
    assert rel_err(loss_s, loss_f) < 1e-6
    assert rel_err(lr_s, lr_f) < 1e-6 and rel_err(lc_s, lc_f) < 1e-6


def test_nce_local_rows_match_full(scp):
    """The multi-GPU layout: every rank evaluates the global loss and back-propagates into its own rows only."""
    N, D, world = 96, 64, 4
    gen = torch.Generator().manual_seed(9)
    a = torch.nn.functional.normalize(torch.randn(N, D, generator=gen), dim=-1).cuda()
    b = torch.nn.functional.normalize(torch.randn(N, D, generator=gen), dim=-1).cuda()
    ids = torch.randint(0, 20, (N,), generator=gen).cuda()
    crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True).cuda()
    af, bf = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    full = crit(af, bf, ids)
    gaf, gbf, gtf = torch.autograd.grad(full, [af, bf, crit.temperature])
    n = N // world
    ga_sum, gt_sum = torch.zeros_like(a), torch.zeros(())
    for r in range(world):
        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        loss_r = crit(ar, br, ids, local_rows=(r * n, (r + 1) * n))
        assert torch.equal(loss_r, full)
        gar, gbr, gtr = torch.autograd.grad(loss_r, [ar, br, crit.temperature])
        assert gar[:r * n].abs().max().item() == 0 if r else True
        assert rel_err(gar[r * n:(r + 1) * n], gaf[r * n:(r + 1) * n]) < 1e-5
        assert rel_err(gbr[r * n:(r + 1) * n], gbf[r * n:(r + 1) * n]) < 1e-5
        ga_sum += gar
        gt_sum = gt_sum + gtr.cpu()
    assert rel_err(ga_sum, gaf) < 1e-5
    assert rel_err(gt_sum, gtf) < 1e-4  # the scale gradient is the sum of the ranks' partial sums


def test_nce_backward_as_first_node_of_a_fresh_autograd_thread(scp):
    """Regression: the backward builds TMA descriptors with a driver call; when it is the FIRST thing PyTorch's autograd
    worker thread ever runs, that thread has no CUDA context bound yet (CUDA_ERROR_INVALID_CONTEXT before the fix)."""
    import subprocess
    import sys
    code = (
        "import sys, torch; sys.path.insert(0, %r); import speechclip_plus_b200 as scp\n"
        "crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True).cuda()\n"
        "a = torch.nn.functional.normalize(torch.randn(64, 128, device='cuda'), dim=-1).requires_grad_(True)\n"
        "b = torch.nn.functional.normalize(torch.randn(64, 128, device='cuda'), dim=-1).requires_grad_(True)\n"
        "loss = crit(a, b, torch.arange(64, device='cuda'))\n"
        "g = torch.autograd.grad(loss, [a, b, crit.temperature])\n"
        "torch.cuda.synchronize(); print('ok', float(g[0].abs().sum()) > 0)\n" % ROOT_DIR)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok True" in r.stdout, r.stderr[-800:]


def test_nce_large_logits_do_not_overflow(scp):
    """The reference exponentiates raw logits (losses.py:232) and overflows for a large learnt scale; LSE does not."""
    N, D = 64, 64
    gen = torch.Generator().manual_seed(2)
    a = torch.nn.functional.normalize(torch.randn(N, D, generator=gen), dim=-1)
    crit = scp.MaskedContrastiveLoss(temperature=1.0 / 200.0, temperature_trainable=False).cuda()
    loss = crit(a.cuda(), a.cuda(), None)
    ref = oracle.nce_forward(a.double(), a.double(), None, 200.0)
    assert torch.isfinite(loss) and rel_err(loss, ref) < TOL


def test_gather_and_compute_loss_golden(scp):
    """N0 + G0 + C0 on one process: un-normalised features -> normalise/pack/(gather) -> hybrid loss (kwClip.py:999-1040)."""
    g = load_golden("hybrid_loss")
    gen = torch.Generator().manual_seed(4)
    # the fixture holds NORMALISED features; scale rows arbitrarily to exercise the normalisation and its backward
    s1 = torch.rand(32, 1, generator=gen) * 3 + 0.5
    s2 = torch.rand(32, 1, generator=gen) * 3 + 0.5
    img = (g["image_feat"] * s1).cuda()
    ca = (g["cascaded_audio_feat"] * s2).cuda().requires_grad_(True)
    pa = (g["parallel_audio_feat"] * s1).cuda().requires_grad_(True)
    crit = scp.MaskedContrastiveLoss(temperature=float(g["temperature"]), temperature_trainable=True).cuda()
    feats = {"id": g["ids"].cuda(), "image_feat": img, "cascaded_audio_feat": ca, "parallel_audio_feat": pa,
             "cif_quantity_out": g["cif_quantity_out"].cuda(), "cif_target_len": g["cif_target_len"].cuda()}
    gathered, rows = scp.gather_loss_feats(feats)
    assert rows == (0, 32)
    assert rel_err(gathered["cascaded_audio_feat"], g["cascaded_audio_feat"]) < 1e-5
    assert torch.equal(gathered["id"].cpu(), g["ids"])
    out = scp.compute_loss(gathered, crit, float(g["cascaded_weight"]), float(g["parallel_weight"]),
                           quantity_loss_weight=float(g["quantity_loss_weight"]),
                           quantity_loss_criteria=torch.nn.L1Loss(), local_rows=rows)
    for key in ("loss", "c_cl_loss", "p_cl_loss", "quantity_loss"):
        assert rel_err(out[key], g[key]) < TOL, key
    gca, gpa, gt = torch.autograd.grad(out["loss"], [ca, pa, crit.temperature])
    assert rel_err(gt, g["grad_temperature"]) < TOL
    # reference gradient w.r.t. the un-normalised features through the oracle (normalise -> loss)
    car = (g["cascaded_audio_feat"] * s2).double().requires_grad_(True)
    par = (g["parallel_audio_feat"] * s1).double().requires_grad_(True)
    ref = oracle.hybrid_loss({"id": g["ids"], "image_feat": g["image_feat"].double(),
                              "cascaded_audio_feat": oracle.l2_normalise(car),
                              "parallel_audio_feat": oracle.l2_normalise(par)},
                             1.0 / float(g["temperature"]), float(g["cascaded_weight"]), float(g["parallel_weight"]))
    rca, rpa = torch.autograd.grad(ref["loss"], [car, par])
    assert norm_err(gca, rca) < TOL and norm_err(gpa, rpa) < TOL


@pytest.mark.parametrize("cfg", [(1024, 512, 128, 3, 1.0, 1.0), (512, 768, 64, 5, 1.5, 0.5)],
                         ids=["c4_hybrid_base_N1024", "c5_hybrid_large_N512"])
def test_hybrid_loss_global_batch_local_rows(scp, cfg):
    """BASELINE configs 4 / 5 as ONE rank of eight sees them: the hybrid loss over the gathered global batch (duplicate
    image ids as in Flickr8k), gradients for this rank's rows only, against the fp64 oracle."""
    N, D, n_local, rank, w_c, w_p = cfg
    gen = torch.Generator().manual_seed(N + D)
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=gen), dim=-1)
    ca = torch.nn.functional.normalize(torch.randn(N, D, generator=gen) + 0.5 * img, dim=-1)
    pa = torch.nn.functional.normalize(torch.randn(N, D, generator=gen) + 0.8 * img, dim=-1)
    ids = torch.randint(0, N // 3, (N,), generator=gen)
    rows = (rank * n_local, (rank + 1) * n_local)
    crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True).cuda()
    cad, pad = ca.cuda().requires_grad_(True), pa.cuda().requires_grad_(True)
    out = scp.compute_loss({"id": ids.cuda(), "image_feat": img.cuda(), "cascaded_audio_feat": cad,
                            "parallel_audio_feat": pad}, crit, w_c, w_p, local_rows=rows)
    gca, gpa, gt = torch.autograd.grad(out["loss"], [cad, pad, crit.temperature])
    car, par = ca.double().requires_grad_(True), pa.double().requires_grad_(True)
    log_scale = torch.tensor(float(crit.temperature.item()), dtype=torch.float64, requires_grad=True)
    ref = oracle.hybrid_loss({"id": ids, "image_feat": img.double(), "cascaded_audio_feat": car,
                              "parallel_audio_feat": par}, log_scale.exp(), w_c, w_p)
    rca, rpa, rt = torch.autograd.grad(ref["loss"], [car, par, log_scale])
    for key in ("loss", "c_cl_loss", "p_cl_loss"):
        assert rel_err(out[key], ref[key]) < TOL, key
    lo, hi = rows
    assert norm_err(gca[lo:hi], rca[lo:hi]) < TOL and norm_err(gpa[lo:hi], rpa[lo:hi]) < TOL
    assert float(gca[:lo].abs().max()) == 0.0 and float(gca[hi:].abs().max()) == 0.0   # other ranks' rows: no gradient here
    # the log-scale gradient is produced in shards as well: the parts of all ranks add up to the reference's gradient
    gt_sum = torch.zeros((), device="cuda")
    for r in range(N // n_local):
        o = scp.compute_loss({"id": ids.cuda(), "image_feat": img.cuda(), "cascaded_audio_feat": cad,
                              "parallel_audio_feat": pad}, crit, w_c, w_p, local_rows=(r * n_local, (r + 1) * n_local))
        gt_sum = gt_sum + torch.autograd.grad(o["loss"], [crit.temperature])[0]
    assert rel_err(gt_sum, rt) < TOL


@pytest.mark.parametrize("name", golden_names("chain_"))
def test_cascaded_chain_golden(scp, name):
    """The cascaded branch from the CLS keywords to the loss, composed as the reference composes it
    (kw_branches.py:390-395 / :744-750 -> kwClip.py:905-907 -> :1021-1025): projection (library Linear) -> N1 batch-norm
    -> fused V1+V3+V4 -> N3 splice + stand-in text tower -> N0 normalise / G0 gather -> S3 loss, against what the
    reference's own GeneralBranch / ClipModel.encode_keywords / MaskedContrastiveLoss produced for the same tensors.
    One backward through the whole chain: every CUDA autograd node hands its gradient to the next."""
    import types
    from speechclip_plus_b200.module.clip_glue import encode_keywords
    g = load_golden(name)
    dynamic = g["keyword_num"].dim() > 0
    B, K, Da = g["audio_feat"].shape
    V, D = g["table"].shape
    clip, _ = _fake_clip(g)
    clip.encode_keywords = types.MethodType(encode_keywords, clip)
    proj = torch.nn.Linear(Da, D).cuda()
    proj.weight.data.copy_(g["proj_weight"]); proj.bias.data.copy_(g["proj_bias"])
    init_bias, init_scale = g["table"].mean(0), g["table"].std(0)                          # kw_branches.py:99-100
    if dynamic:
        bn = scp.Kw_BatchNorm_dynamic(kw_dim=D, init_bias=init_bias, init_scale=init_scale, std_scale=1, learnable=True)
    else:
        bn = scp.Kw_BatchNorm(kw_num=K, kw_dim=D, batchnorm_type="eachKw", init_bias=init_bias, init_scale=init_scale,
                              std_scale=1, learnable=True, parallel=True)
    state_in = {k[len("bn_in__"):].replace("__", "."): v for k, v in g.items() if k.startswith("bn_in__")}
    # the constructor's own initialisation from the table statistics equals the reference's (kw_bn.py:68-95)
    assert rel_err(bn.state_dict()["bn_layer.weight"], state_in["bn_layer.weight"]) < 1e-6
    assert rel_err(bn.state_dict()["bn_layer.bias"], state_in["bn_layer.bias"]) < 1e-6
    bn.load_state_dict(state_in)                                                           # same keys as the reference
    bn = bn.cuda().train()
    branch = types.SimpleNamespace(clip=clip, vector_quantizer=scp.SimpleVectorQuantizer("fixed=0.1").cuda().train(),
                                   project_feats_to_CLIPspace=lambda f: bn(proj(f)))      # kw_branches.py:143-156
    crit = scp.MaskedContrastiveLoss(temperature=float(g["temperature"]), temperature_trainable=True).cuda()

    audio_feat = g["audio_feat"].cuda().requires_grad_(True)
    vq_results, keywords = scp.fused_vq_audio_features(branch, audio_feat)
    keyword_num = g["keyword_num"].cuda() if dynamic else int(g["keyword_num"])
    cascaded = clip.encode_keywords(keywords, keyword_num)
    gathered, rows = scp.gather_loss_feats({"id": g["ids"].cuda(), "image_feat": g["image_feat"].cuda(),
                                            "cascaded_audio_feat": cascaded})
    out = scp.compute_loss(gathered, crit, 1.0, 0.0, local_rows=rows)

    assert torch.equal(vq_results["targets"].cpu(), g["targets"])                          # code indices: bit-exact
    assert rel_err(keywords, g["keywords"]) < TOL
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t"):
        assert rel_err(vq_results[key], g[key]) < TOL, key
    for key in ("running_mean", "running_var"):
        assert rel_err(bn.state_dict()[f"bn_layer.{key}"], g[f"bn_out__bn_layer__{key}"]) < TOL, key
    assert int(bn.state_dict()["bn_layer.num_batches_tracked"]) == int(g["bn_out__bn_layer__num_batches_tracked"])
    assert rel_err(cascaded, g["cascaded_audio_feat"]) < TOL
    assert rel_err(out["loss"], g["loss"]) < TOL and out["loss"] is out["c_cl_loss"]
    params = {"audio_feat": audio_feat, "proj_weight": proj.weight, "proj_bias": proj.bias,
              "bn_weight": bn.bn_layer.weight, "bn_bias": bn.bn_layer.bias,
              "mix_w": clip.model.transformer.mix.weight, "mix_b": clip.model.transformer.mix.bias,
              "ln_weight": clip.model.ln_final.weight, "ln_bias": clip.model.ln_final.bias,
              "temperature": crit.temperature}
    grads = torch.autograd.grad(out["loss"], list(params.values()))
    for k, gr in zip(params, grads):
        assert norm_err(gr, g[f"grad_{k}"]) < TOL, k


@pytest.mark.parametrize("which", ["reference_head64", "flickr_sized_8112"])
def test_reduced_vocab_chain(scp, which, tmp_path):
    """N2 end to end on the GPU (clip_official.py:63-108 -> kw_branches.py:158-197 -> clip_official.py:222-279): the
    by-frequency usage table -> reduce_subword_embedding (through the installed ClipModel.__init__ wrapper) -> the fused VQ
    against the REDUCED table with the quantiser's default prob_msk = [0, 2, 3] = pad / SOT / EOT of that table ->
    encode_keywords spliced with the reduced SOT / EOT ids -> loss; every stage against the oracle on the reduced table.
    `reference_head64`: the first 64 rows of the reference's own avssl/data/flickr_stat/text_clip_vocab_usage_byfreq.npy;
    `flickr_sized_8112`: a synthetic table of the Flickr recipe's size."""
    import types
    import numpy as np
    from speechclip_plus_b200.model import kwclip_glue
    from speechclip_plus_b200.module.clip_glue import encode_keywords
    V_full, D, B, K, L = 49408, 512, 16, 8, 77
    gen = torch.Generator().manual_seed(99)
    if which == "reference_head64":
        usage = np.load(os.path.join(ROOT_DIR, "tests", "golden", "vocab_usage_byfreq_head64.npy"))
    else:
        rest = torch.randperm(V_full - 2, generator=gen)[:8112 - 4].numpy() + 1      # ids 1 .. 49406 (0 / SOT / EOT excluded)
        rest = rest[(rest != 49406)][:8112 - 4]
        ids = np.concatenate([[0, rest[0], 49406, 49407], rest[1:]])
        counts = np.sort(torch.randint(1, 10 ** 6, (len(ids),), generator=gen).numpy())[::-1]
        usage = np.stack([ids, counts], 1).astype(np.int64)
    path = tmp_path / "usage.npy"
    np.save(path, usage)
    full = torch.randn(V_full, D, generator=gen) * 0.02

    def base_init(self, name, device="cpu", image_encoder_trainable=False, text_encoder_trainable=False,
                  reduce_subword_embbedding=None, **kw):
        torch.nn.Module.__init__(self)
        emb = torch.nn.Embedding.from_pretrained(full.clone()).cuda()

        class Tower(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.mix = torch.nn.Linear(D, D)

            def forward(self, x):  # (L, N, D): mixes the positions, like attention does, so that the EOT feature
                return torch.tanh(self.mix(x) + x.mean(dim=0, keepdim=True))  # depends on the keyword positions

        torch.manual_seed(3)
        self.model = types.SimpleNamespace(token_embedding=emb, positional_embedding=(torch.randn(L, D) * 0.01).cuda(),
                                           transformer=Tower().cuda(), ln_final=torch.nn.LayerNorm(D).cuda(),
                                           text_projection=(torch.randn(D, D) * D ** -0.5).cuda())
        self.text_encoder_trainable = text_encoder_trainable
        self.tokenizer = types.SimpleNamespace(encoder={"<|startoftext|>": 49406, "<|endoftext|>": 49407})
        self.selected_text_emb_ids = None
        self.device = torch.device("cuda")

    cls = type("ClipModelStandIn", (torch.nn.Module,), {"__init__": kwclip_glue.clipmodel_init(base_init),
                                                        "encode_keywords": encode_keywords})
    clip = cls("ViT-B/32", reduce_subword_embbedding=str(path))
    Vr = len(usage)
    table = clip.model.token_embedding.weight
    assert table.shape == (Vr, D) and not table.requires_grad
    assert torch.equal(table.cpu(), full[torch.from_numpy(usage[:, 0].copy())])
    assert (clip.startOfTxt_reduced, clip.endOfTxt_reduced) == (2, 3) and clip.original2Reduced[0] == 0
    # keywords near table rows, some of them near the masked pad / SOT / EOT rows (which must never be chosen)
    pick = torch.randint(0, Vr, (B, K), generator=gen)
    pick[:, 0] = torch.tensor([0, 2, 3] * B)[:B]
    kw = (table.cpu()[pick] * 2.0 + 0.004 * torch.randn(B, K, D, generator=gen)).cuda().requires_grad_(True)
    vq = scp.SimpleVectorQuantizer("fixed=0.1").cuda().train()
    branch = types.SimpleNamespace(clip=clip, vector_quantizer=vq, project_feats_to_CLIPspace=lambda f: f)
    res, keywords = scp.fused_vq_audio_features(branch, kw)                 # default prob_msk = [0, 2, 3]
    ref, kw_ref = oracle.vq_audio_features(kw.detach().double().cpu(), table.double().cpu(), torch.tensor([0.1], dtype=torch.float64))
    assert torch.equal(res["targets"].cpu(), ref["targets"])
    assert not bool(((res["targets"] == 0) | (res["targets"] == 2) | (res["targets"] == 3)).any())
    assert rel_err(keywords, kw_ref) < TOL
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t"):
        assert rel_err(res[key], ref[key]) < TOL, key
    out = clip.encode_keywords(keywords, K)
    x_ref, eot_idx = oracle.splice_keywords(kw_ref.float(), K, table.cpu(), clip.model.positional_embedding.cpu(), 2, 3)
    with torch.no_grad():
        t = clip.model
        y = t.ln_final(t.transformer(x_ref.cuda().permute(1, 0, 2)).permute(1, 0, 2))
        out_ref = y[torch.arange(B), eot_idx.cuda()] @ t.text_projection
    assert rel_err(out, out_ref) < TOL
    crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True).cuda()
    img = torch.randn(B, D, generator=gen).cuda()
    ids = torch.randint(0, 5, (B,), generator=gen).cuda()
    gathered, rows = scp.gather_loss_feats({"id": ids, "image_feat": img, "cascaded_audio_feat": out})
    loss = scp.compute_loss(gathered, crit, 1.0, 0.0)["loss"]
    loss_ref = oracle.nce_forward(oracle.l2_normalise(out_ref.double().cpu()), oracle.l2_normalise(img.double().cpu()),
                                  ids.cpu(), 1 / 0.07)
    assert rel_err(loss, loss_ref) < TOL
    (g_kw,) = torch.autograd.grad(loss, [kw])
    assert torch.isfinite(g_kw).all() and g_kw.abs().sum() > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_vq_under_dataparallel_replicas(scp):
    """The reference trains with strategy: dp (SURVEY section 8(b) "Threading"): nn.DataParallel replicas are shallow
    copies that share the module's table cache and run concurrently, one thread per GPU.  Every replica must use ITS
    device's prepared table (per-device cache entries behind a lock), and the gathered result must equal the
    single-device one."""
    B, K, V, D = 32, 8, 8112, 512
    gen = torch.Generator().manual_seed(17)
    table = (torch.randn(V, D, generator=gen) * 0.02)
    kw = torch.randn(B, K, D, generator=gen) * 0.02

    class Wrap(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vq = scp.SimpleVectorQuantizer("fixed=0.1")
            self.table = torch.nn.Parameter(table.clone(), requires_grad=False)

        def forward(self, x):
            res, out = self.vq.quantize_keywords(x, self.table)
            return out, res["targets"]

    m = Wrap().cuda(0).train()
    single_out, single_idx = m(kw.cuda(0))
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    for _ in range(3):  # repeated calls: the cache entries of both devices stay valid side by side
        out, idx = dp(kw.cuda(0))
        assert torch.equal(idx.cpu(), single_idx.cpu())
        assert torch.equal(out.cpu(), single_out.cpu())
    assert len(m.vq._table_cache._entries) == 2


def test_install_patches_reference_namespaces(scp):
    import sys
    import types
    pkg = "fake_avssl"
    mods = {}
    for name in [pkg, f"{pkg}.module", f"{pkg}.module.losses", f"{pkg}.module.weighted_sum",
                 f"{pkg}.module.speech_encoder_plus", f"{pkg}.module.speechclip_c_modules",
                 f"{pkg}.module.speechclip_c_modules.my_vector_quantizer",
                 f"{pkg}.module.speechclip_c_modules.vector_quantizers", f"{pkg}.module.speechclip_c_modules.kw_bn",
                 f"{pkg}.module.clip_official", f"{pkg}.module.cif", f"{pkg}.util", f"{pkg}.util.data_utils",
                 f"{pkg}.model", f"{pkg}.model.kw_branches", f"{pkg}.model.kwClip"]:
        mods[name] = types.ModuleType(name)
        sys.modules[name] = mods[name]
    # the classes whose methods install() wraps (model/kwclip_glue.py)
    mods[f"{pkg}.model.kwClip"].KWClip_GeneralTransformer = type(
        "KWClip_GeneralTransformer", (), {"compute_loss": lambda self, d: None, "forward": lambda self, b: None})
    for cls_name in ("FairseqSpeechEncoder_Hubert", "S3prlSpeechEncoderPlus"):
        setattr(mods[f"{pkg}.module.speech_encoder_plus"], cls_name, type(cls_name, (), {"forward": lambda self, wav: wav}))

    class GeneralBranch:  # noqa: D401
        pass

    mods[f"{pkg}.model.kw_branches"].GeneralBranch = GeneralBranch
    mods[f"{pkg}.module.clip_official"].ClipModel = type("ClipModel", (), {})
    try:
        done = scp.install(pkg, strict=True)
        assert all(done.values())
        assert mods[f"{pkg}.module.losses"].MaskedContrastiveLoss is scp.MaskedContrastiveLoss
        assert mods[f"{pkg}.model.kw_branches"].Kw_BatchNorm is scp.Kw_BatchNorm
        assert mods[f"{pkg}.model.kw_branches"].CIF is scp.CIF and mods[f"{pkg}.module.cif"].CIF is scp.CIF
        from speechclip_plus_b200.module import clip_glue
        assert mods[f"{pkg}.module.clip_official"].ClipModel.encode_keywords is clip_glue.encode_keywords
        assert mods[f"{pkg}.model.kw_branches"].get_keypadding_mask is clip_glue.get_keypadding_mask
        assert mods[f"{pkg}.module.speechclip_c_modules.kw_bn"].Kw_BatchNorm_dynamic is scp.Kw_BatchNorm_dynamic
        assert getattr(mods[f"{pkg}.module.speechclip_c_modules.vector_quantizers"], "SimpleVectorQuantizer") \
            is scp.SimpleVectorQuantizer
        assert mods[f"{pkg}.model.kwClip"].KWClip_GeneralTransformer.compute_loss._scp_installed
        assert mods[f"{pkg}.module.speech_encoder_plus"].S3prlSpeechEncoderPlus.forward._scp_installed
        # the patched method runs the fused path on a duck-typed branch (projection = identity)
        V, D = 512, 64
        table = torch.randn(V, D).cuda() * 0.02
        emb = torch.nn.Embedding(V, D).cuda()
        emb.weight.data.copy_(table)
        emb.weight.requires_grad_(False)
        br = GeneralBranch()
        br.project_feats_to_CLIPspace = lambda x: x
        br.clip = types.SimpleNamespace(model=types.SimpleNamespace(token_embedding=emb))
        br.vector_quantizer = scp.SimpleVectorQuantizer("fixed=0.1").cuda()
        res, kws = br.vq_audio_features(torch.randn(2, 3, D).cuda())
        assert kws.shape == (2, 3, D) and res["targets"].shape == (2, 3, 1)
    finally:
        for name in mods:
            sys.modules.pop(name, None)
