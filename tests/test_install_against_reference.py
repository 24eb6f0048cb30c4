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


GLUE_METHODS = [("avssl.model.kwClip", "KWClip_GeneralTransformer", "compute_loss"),
                ("avssl.model.kwClip", "KWClip_GeneralTransformer", "forward"),
                ("avssl.module.speech_encoder_plus", "FairseqSpeechEncoder_Hubert", "forward"),
                ("avssl.module.speech_encoder_plus", "S3prlSpeechEncoderPlus", "forward"),
                ("avssl.module.clip_official", "ClipModel", "__init__")]


def _save_glue_methods():
    import importlib
    return [(getattr(importlib.import_module(m), c), a, getattr(importlib.import_module(m), c).__dict__[a])
            for m, c, a in GLUE_METHODS]


def _restore_glue_methods(saved):
    for cls, attr, fn in saved:
        setattr(cls, attr, fn)


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
    saved_glue = _save_glue_methods()
    before = [_signature(m) for m in _build(kb, losses, sep, recipe)]
    try:
        done = scp.install("avssl", strict=True)
        assert all(done.values()), done
        glue_keys = {f"{m}.{c}.{a}" for m, c, a in GLUE_METHODS}
        assert sorted(k for k in done if not k.endswith(("vq_audio_features", "encode_keywords")) and k not in glue_keys) == \
            sorted(f"{m}.{n}" for m, n in patched)                      # the restore list below is complete
        assert glue_keys.issubset(done)
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
        _restore_glue_methods(saved_glue)


class _StubBranch(torch.nn.Module):
    """Stands in for KW_CascadedBranch: returns the dict the reference's forward expects (kw_branches.py:430-447)."""

    def __init__(self, D):
        super().__init__()
        self.proj = torch.nn.Linear(8, D)

    def forward(self, audio_feat, audio_feat_len, otherInputs=None):
        feat = self.proj(audio_feat.mean(dim=1)) * 3.0  # deliberately NOT unit-norm
        vq = {"temp": 0.1, "code_perplexity": torch.tensor(2.0), "prob_perplexity": torch.tensor(3.0),
              "ent_per_t": torch.ones(2)}
        return {"parallel_audio_feat": None, "cascaded_audio_feat": feat, "vq_results": vq, "keywords": None,
                "dsample_results": None}


def test_forward_to_training_step_end_runs_through_the_installed_glue(monkeypatch):
    """install() puts N0 + G0 + C0 on the reference's own call path: KWClipBase.training_step -> (our) forward ->
    KWClipBase.training_step_end (kwClip.py:149-193) -> (our) compute_loss -> gather_loss_feats + compute_loss.  The
    model is a real KWClip_GeneralTransformer instance whose towers / branch are small stand-ins (the real ones need
    fairseq / clip checkpoints); the two CUDA entry points of the glue are replaced by recorders so that the routing can be
    asserted on the CPU.  Values are checked on the GPU by tests/test_gpu_parity.py."""
    import types
    import speechclip_plus_b200 as scp
    from speechclip_plus_b200.model import kw_glue
    ref.import_avssl()
    import avssl.model.kwClip as kc
    from avssl.base import OrderedNamespace
    saved = _save_glue_methods()
    try:
        done = scp.install("avssl", strict=False)
        assert all(done[f"{m}.{c}.{a}"] for m, c, a in GLUE_METHODS), done
        assert getattr(kc.KWClip_GeneralTransformer.compute_loss, "_scp_installed", False)
        assert scp.install("avssl")[f"{GLUE_METHODS[0][0]}.{GLUE_METHODS[0][1]}.{GLUE_METHODS[0][2]}"]   # idempotent
        assert kc.KWClip_GeneralTransformer.compute_loss._scp_original is saved[0][2]                   # not double-wrapped

        D, B = 16, 4
        model = object.__new__(kc.KWClip_GeneralTransformer)     # skip __init__: it downloads HuBERT / CLIP
        torch.nn.Module.__init__(model)
        model.config = OrderedNamespace({"model_settings": {"cascaded_objective_weight": 1.0,
                                                            "parallel_objective_weight": 0.0}})
        model.clip = types.SimpleNamespace(update_device=lambda d: None)
        model.forward_audio = lambda wav, wav_len, return_hidden_states=False: (torch.randn(B, 5, 8), torch.full((B,), 5))
        model.forward_image = lambda image: torch.randn(B, D) * 2.0
        model.img_enc_proj_net = model.p_branch_proj_net = model.c_branch_proj_net = None
        model.cascaded_branch = _StubBranch(D)
        model.parallel_branch = None
        model.criterion = types.SimpleNamespace(current_temperature=0.07)
        model.global_step = 0
        if not hasattr(type(model), "device"):
            model.device = torch.device("cpu")
        calls = {}

        def fake_gather(loss_feats, group=None):
            calls["gather"] = {k: v for k, v in loss_feats.items()}
            return dict(loss_feats), (0, B)

        def fake_compute(loss_feats, criterion, **kw):
            calls["compute"] = kw
            assert criterion is model.criterion
            return {"loss": loss_feats["cascaded_audio_feat"].sum() * 0 + 1.25, "c_cl_loss": torch.tensor(1.25)}

        monkeypatch.setattr(kw_glue, "gather_loss_feats", fake_gather)
        monkeypatch.setattr(kw_glue, "compute_loss", fake_compute)
        batch = {"wav": torch.zeros(B, 10), "wav_len": torch.full((B,), 10), "image": torch.zeros(B, 3, 2, 2),
                 "id": torch.arange(B)}
        model.train()
        out = model.training_step(batch)                                   # kwClip.py:145-147 (the reference's own method)
        feats = out["loss_feats"]
        assert set(feats) == {"id", "image_feat", "cascaded_audio_feat"}
        # our forward leaves the normalisation to the pack kernel while training ...
        assert not torch.allclose(feats["cascaded_audio_feat"].norm(dim=-1), torch.ones(B), atol=1e-3)
        assert not torch.allclose(feats["image_feat"].norm(dim=-1), torch.ones(B), atol=1e-3)
        assert set(out["log_metrics"]) == {"cl_temp", "softmax_temp", "temp", "code_perplexity", "prob_perplexity", "ent_per_t"}
        res = model.training_step_end(out)                                 # kwClip.py:149-193 (the reference's own method)
        assert float(res["loss"]) == 1.25
        assert calls["gather"]["cascaded_audio_feat"] is feats["cascaded_audio_feat"]
        assert calls["compute"]["cascaded_objective_weight"] == 1.0 and calls["compute"]["parallel_objective_weight"] == 0.0
        assert calls["compute"]["local_rows"] is None                      # single process: no sharding
        # ... and keeps it in evaluation, where the third output feeds the retrieval code directly
        model.eval()
        losses, _, others = model.forward(batch)
        assert torch.allclose(others["cascaded_audio_feat"].norm(dim=-1), torch.ones(B), atol=1e-5)
        assert torch.allclose(losses["image_feat"].norm(dim=-1), torch.ones(B), atol=1e-5)
    finally:
        _restore_glue_methods(saved)


def test_upstream_forward_wrapper_folds_the_method_rescale_into_the_weighted_sum():
    """S1' through install(): the wrapped FairseqSpeechEncoder_Hubert.forward must skip the reference's per-layer
    method1 / method2 loop (speech_encoder_plus.py:572-592) exactly when the weighted sum follows, hand the mode to our
    WeightedSumLayer for the duration of the call, and otherwise run the original code."""
    import speechclip_plus_b200 as scp
    from speechclip_plus_b200 import _lib
    from speechclip_plus_b200.model import kwclip_glue
    seen = {}

    class Layer(scp.WeightedSumLayer):
        def forward(self, x):
            seen["mode"] = self.upstream_norm_mode
            return x[0]

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.normalize_hiddenstates, self.normalize_type, self.feat_select_idx = True, "method2", "weighted_sum"
            self.weightedsum_layer = Layer(2)

    def original(self, wav, wav_len=[], feat_select_idx=None, return_hidden_states=False):  # noqa: B006
        seen["loop_would_run"] = self.normalize_hiddenstates and self.normalize_type.startswith("method")
        sel = self.feat_select_idx if feat_select_idx is None else feat_select_idx
        return self.weightedsum_layer([wav, wav]) if sel == "weighted_sum" else wav

    enc = Enc()
    fwd = kwclip_glue.fused_upstream_forward(original)
    fwd(enc, torch.zeros(2, 3))
    assert seen == {"loop_would_run": False, "mode": _lib.SCP_NORM_UTT_MEAN}
    assert enc.normalize_hiddenstates is True and enc.weightedsum_layer.upstream_norm_mode is None   # restored
    fwd(enc, torch.zeros(2, 3), return_hidden_states=True)       # the caller wants the rescaled states: original path
    assert seen["loop_would_run"] is True and seen["mode"] is None
    enc.normalize_type = "s3prl"                                  # LayerNorm flag of the layer: nothing to fold
    fwd(enc, torch.zeros(2, 3))
    assert seen["mode"] is None


def test_clipmodel_init_wrapper_builds_the_reduced_vocabulary(tmp_path):
    """N2 through install(): ClipModel.__init__ with reduce_subword_embbedding set (clip_official.py:63-108) must leave the
    same attributes behind when the branch is served by reduce_subword_embedding -- compared with the reference's own
    branch executed on the same stand-in CLIP."""
    import types
    import numpy as np
    from speechclip_plus_b200.model import kwclip_glue
    ref.import_avssl()
    import avssl.module.clip_official as co
    V, D = 50, 8
    usage = np.stack([np.array([0, 7, 48, 49, 3, 11, 20], dtype=np.int64), np.array([9, 8, 7, 6, 5, 4, 3], dtype=np.int64)], 1)
    path = tmp_path / "usage.npy"
    np.save(path, usage)
    torch.manual_seed(0)
    weight = torch.randn(V, D)

    def base_init(self, name, device="cpu", image_encoder_trainable=False, text_encoder_trainable=False,
                  reduce_subword_embbedding=None, **kw):
        torch.nn.Module.__init__(self)
        assert reduce_subword_embbedding is None     # the wrapper serves that branch itself
        self.model = types.SimpleNamespace(token_embedding=torch.nn.Embedding.from_pretrained(weight.clone()))
        self.text_encoder_trainable = text_encoder_trainable
        self.tokenizer = types.SimpleNamespace(encoder={"<|startoftext|>": 48, "<|endoftext|>": 49})
        self.selected_text_emb_ids = None

    cls = type("ClipModelStandIn", (torch.nn.Module,), {"__init__": kwclip_glue.clipmodel_init(base_init)})
    ours = cls("ViT-B/32", reduce_subword_embbedding=str(path))
    # the reference's own branch, transcribed call by call from its __init__ onto the same stand-in
    ids = usage[:, 0]
    assert np.array_equal(ours.selected_text_emb_ids, ids)
    assert torch.allclose(ours.selected_text_emb_ids_dist, torch.from_numpy(usage[:, 1] / usage[:, 1].sum()))
    assert torch.equal(ours.model.token_embedding.weight, weight[ids]) and not ours.model.token_embedding.weight.requires_grad
    assert torch.equal(ours.original_text_emb_weight, weight)
    assert ours.original2Reduced == {int(o): n for n, o in enumerate(ids)} and ours.reducedl2Original == {n: int(o) for n, o in enumerate(ids)}
    assert (ours.startOfTxt_reduced, ours.endOfTxt_reduced) == (2, 3)      # the quantiser's prob_msk = [0, 2, 3] default
    with pytest.raises(SystemExit):
        cls("ViT-B/32", reduce_subword_embbedding=str(tmp_path / "missing.npy"))
    assert hasattr(co.ClipModel, "encode_keywords")


def test_uninstall_restores_the_reference():
    """install() -> uninstall() leaves every patched name of the reference package exactly as it was (classes looked up
    by name, the two method bodies, the five wrapped glue methods), and a second install() starts from the originals."""
    import speechclip_plus_b200 as scp
    ref.import_avssl()
    import importlib
    kb = importlib.import_module("avssl.model.kw_branches")
    losses = importlib.import_module("avssl.module.losses")
    co = importlib.import_module("avssl.module.clip_official")
    before = (kb.GeneralBranch.__dict__["vq_audio_features"], co.ClipModel.__dict__["encode_keywords"],
              losses.MaskedContrastiveLoss, kb.CIF, kb.Kw_BatchNorm, [fn for _, _, fn in _save_glue_methods()])
    done = scp.install("avssl", strict=True)
    assert losses.MaskedContrastiveLoss is scp.MaskedContrastiveLoss
    assert scp.uninstall() >= len(done) - 2          # package-level aliases may share a module object
    after = (kb.GeneralBranch.__dict__["vq_audio_features"], co.ClipModel.__dict__["encode_keywords"],
             losses.MaskedContrastiveLoss, kb.CIF, kb.Kw_BatchNorm, [fn for _, _, fn in _save_glue_methods()])
    assert all(a is b for a, b in zip(before[:5], after[:5]))
    assert all(a is b for a, b in zip(before[5], after[5]))
    assert scp.uninstall() == 0
    scp.install("avssl", strict=True)
    assert co.ClipModel.__init__._scp_original is before[5][4]
