"""CPU, build container only (skipped where /root/reference is absent, i.e. on the GPU box): `install()` against the REAL
reference package.  The reference's own constructors -- driven by its own shipped YAML recipes -- must build OUR classes
through its plugin points (kwClip.py:84, kw_branches.py:75-91, :95, :629, :619; speech_encoder_plus.py:24) and end up with
exactly the parameter / buffer names and shapes of the unpatched model, so that released checkpoints load unchanged
("state_dict / checkpoint compatibility").  No compute is executed: the modules are only built.
"""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _ref_import as ref  # noqa: E402

pytestmark = pytest.mark.skipif(not ref.reference_available(), reason="needs the reference checkout (/root/reference)")

RECIPES = [
    # YAML (relative to the reference root), branch class, constructor takes out_dim (the hybrid branches)
    ("config/speechCLIP/model_base/spchclp_c.yaml", "KW_CascadedBranch", False),
    ("config/speechCLIP+/model_base/spchclip_c+.yaml", "KW_CascadedBranchPlus", False),
    ("config/speechCLIP+/model_base/spchclip_h.yaml", "KW_HybridBranch", True),
    ("config/speechCLIP+/model_base/spchclip_h+.yaml", "KW_HybridBranchPlus", True),
]


def _fake_clip(V=600, D=512):
    emb = torch.nn.Embedding(V, D)
    emb.weight.requires_grad_(False)
    model = types.SimpleNamespace(token_embedding=emb, positional_embedding=torch.zeros(77, D),
                                  transformer=torch.nn.Identity(), ln_final=torch.nn.LayerNorm(D),
                                  text_projection=torch.zeros(D, D))
    return types.SimpleNamespace(model=model, device=torch.device("cpu"), selected_text_emb_ids=None,
                                 tokenizer=types.SimpleNamespace(encoder={"<|startoftext|>": V - 2, "<|endoftext|>": V - 1}))


def _build(kb, losses, wsum_mod, recipe):
    import yaml
    from avssl.base import OrderedNamespace
    path, cls_name, hybrid = recipe
    cfg = OrderedNamespace(yaml.safe_load(open(os.path.join(ref.REF_ROOT, path))))
    torch.manual_seed(0)
    args = (cfg, 768, 512, 512, _fake_clip()) if hybrid else (cfg, 768, 512, _fake_clip())
    branch = getattr(kb, cls_name)(*args)
    criterion = getattr(losses, cfg.cl_loss.type)(**cfg.cl_loss.args)                 # kwClip.py:84
    wsum = wsum_mod.WeightedSumLayer(n_weights=13, normalize_features=False)          # speech_encoder_plus.py:218-220
    return branch, criterion, wsum


def _signature(module):
    sd = module.state_dict()
    return {k: (tuple(v.shape), v.dtype) for k, v in sd.items()}, sorted(n for n, _ in module.named_parameters())


@pytest.mark.parametrize("recipe", RECIPES, ids=[r[1] for r in RECIPES])
def test_reference_constructors_build_our_classes_with_identical_state(recipe):
    import speechclip_plus_b200 as scp
    ref.import_avssl()
    import avssl.model.kw_branches as kb
    import avssl.module.losses as losses
    import avssl.module.speech_encoder_plus as sep

    import importlib
    patched = [("avssl.module.losses", "MaskedContrastiveLoss"), ("avssl.module", "MaskedContrastiveLoss"),
               ("avssl.module.weighted_sum", "WeightedSumLayer"), ("avssl.module.speech_encoder_plus", "WeightedSumLayer"),
               ("avssl.module", "WeightedSumLayer"),
               ("avssl.module.speechclip_c_modules.my_vector_quantizer", "SimpleVectorQuantizer"),
               ("avssl.module.speechclip_c_modules.vector_quantizers", "SimpleVectorQuantizer"),
               ("avssl.module.speechclip_c_modules.kw_bn", "Kw_BatchNorm"), ("avssl.model.kw_branches", "Kw_BatchNorm"),
               ("avssl.module.speechclip_c_modules.kw_bn", "Kw_BatchNorm_dynamic"),
               ("avssl.model.kw_branches", "Kw_BatchNorm_dynamic"), ("avssl.module.cif", "CIF"),
               ("avssl.model.kw_branches", "CIF"), ("avssl.util.data_utils", "get_keypadding_mask"),
               ("avssl.model.kw_branches", "get_keypadding_mask")]
    saved = [(importlib.import_module(m), n, getattr(importlib.import_module(m), n)) for m, n in patched]
    clip_model = importlib.import_module("avssl.module.clip_official").ClipModel
    saved_methods = (kb.GeneralBranch.vq_audio_features, clip_model.encode_keywords)
    before = [_signature(m) for m in _build(kb, losses, sep, recipe)]
    try:
        done = scp.install("avssl", strict=True)
        assert all(done.values()), done
        assert sorted(k for k in done if not k.endswith(("vq_audio_features", "encode_keywords"))) == \
            sorted(f"{m}.{n}" for m, n in patched)                      # the restore list below is complete
        branch, criterion, wsum = _build(kb, losses, sep, recipe)
        # the reference's own plumbing instantiated OUR classes
        assert type(branch.vector_quantizer) is scp.SimpleVectorQuantizer
        assert type(criterion) is scp.MaskedContrastiveLoss and type(wsum) is scp.WeightedSumLayer
        if hasattr(branch, "bn_layer"):
            assert type(branch.bn_layer) in (scp.Kw_BatchNorm, scp.Kw_BatchNorm_dynamic)
        if hasattr(branch, "downsampling"):
            assert type(branch.downsampling) is scp.CIF
        assert kb.GeneralBranch.vq_audio_features is not saved_methods[0]
        # ... with the same state_dict keys / shapes / dtypes and the same trainable-parameter names
        after = [_signature(m) for m in (branch, criterion, wsum)]
        for (sd0, p0), (sd1, p1), what in zip(before, after, ("branch", "criterion", "weighted sum")):
            assert sd0 == sd1, (what, set(sd0) ^ set(sd1))
            assert p0 == p1, what
    finally:  # leave the reference package as it was found
        for m, n, v in saved:
            setattr(m, n, v)
        kb.GeneralBranch.vq_audio_features = saved_methods[0]
        clip_model.encode_keywords = saved_methods[1]
