"""CPU: host-side logic of the drop-in modules (constructor semantics, state_dict keys, error behaviour)."""
import math

import pytest
import torch

import speechclip_plus_b200 as scp
from speechclip_plus_b200.model import kw_glue
from speechclip_plus_b200.module.vector_quantizers import _masked_array


def test_weighted_sum_layer_contract():
    layer = scp.WeightedSumLayer(n_weights=13, normalize_features=True)
    assert layer.n_weights == 13 and layer.normalize_features
    assert list(layer.state_dict().keys()) == ["weights"]                      # weighted_sum.py:21
    assert layer.weights.shape == (13,) and layer.weights.dtype == torch.float32
    assert torch.count_nonzero(layer.weights) == 0
    with pytest.raises(AssertionError):
        layer([torch.zeros(1, 2, 4)] * 12)                                       # weighted_sum.py:36
    with pytest.raises(scp.ScpError):
        layer([torch.zeros(1, 2, 4)] * 13)                                       # CPU tensors: no fallback
    with pytest.raises(scp.ScpError):
        scp.WeightedSumLayer(n_weights=64)


def test_upstream_tail_host_logic():
    from speechclip_plus_b200.module.speech_encoder_plus import fuse_upstream_features, upstream_feat_len
    # feat_len: Python round (half to even) then clamp to T (speech_encoder_plus.py:604-611)
    fl = upstream_feat_len([800, 1120, 2400, 160, 99999], 320, 9)
    assert fl.dtype == torch.int64 and fl.tolist() == [2, 4, 8, 0, 9]
    assert scp.WeightedSumLayer(4, True, normalize_type="method1").normalize_type == "method1"
    with pytest.raises(AssertionError):
        scp.WeightedSumLayer(4, True, normalize_type="zscore")               # speech_encoder_plus.py:377
    ln_layer = scp.WeightedSumLayer(3, normalize_features=True)
    with pytest.raises(scp.ScpError):                                          # method* excludes the LN flag (:472-476)
        fuse_upstream_features([torch.zeros(1, 2, 4)] * 3, ln_layer, True, "method2")
    with pytest.raises(AssertionError):
        fuse_upstream_features([torch.zeros(1, 2, 4)] * 2, ln_layer, False, "s3prl")
    with pytest.raises(scp.ScpError):                                          # CPU tensors: no fallback
        fuse_upstream_features([torch.zeros(1, 2, 4)] * 3, scp.WeightedSumLayer(3), True, "method1")


def test_vocabulary_reduction_matches_reference_format():
    """N2: the (V',2) [token id, count] by-frequency table of the reference (first 64 rows of
    avssl/data/flickr_stat/text_clip_vocab_usage_byfreq.npy) -> reduced table, id maps, SOT / EOT positions."""
    import os
    import numpy as np
    from speechclip_plus_b200.module.clip_glue import ReducedVocab, reduce_subword_embedding
    usage = np.load(os.path.join(os.path.dirname(__file__), "golden", "vocab_usage_byfreq_head64.npy"))
    assert usage.shape == (64, 2) and usage.dtype == np.int64
    emb = torch.nn.Embedding(49408, 8)
    reduced, vocab, original = reduce_subword_embedding(emb, usage, sot_token=49406, eot_token=49407)
    assert len(vocab) == 64 and reduced.weight.shape == (64, 8) and not reduced.weight.requires_grad
    assert original is emb.weight
    assert torch.equal(reduced.weight, emb.weight.detach()[torch.from_numpy(usage[:, 0])])   # clip_official.py:84-86
    # rows 0 / 2 / 3 are pad, SOT, EOT: exactly the quantiser's default prob_msk (my_vector_quantizer.py:64)
    assert vocab.reducedl2Original[0] == 0 and vocab.startOfTxt_reduced == 2 and vocab.endOfTxt_reduced == 3
    assert vocab.original2Reduced[320] == 1 and vocab.reducedl2Original[1] == 320
    assert abs(float(vocab.selected_text_emb_ids_dist.sum()) - 1.0) < 1e-12
    assert vocab.selected_text_emb_ids_dist.dtype == torch.float64
    with pytest.raises(KeyError):                                           # SOT missing from the table (:103-105)
        ReducedVocab(usage[4:], 49406, 49407)
    with pytest.raises(ValueError):
        ReducedVocab(usage[:, 0], 49406, 49407)


def test_vector_quantizer_constructor_semantics():
    fixed = scp.SimpleVectorQuantizer("fixed=0.1")
    assert fixed.temp_type == "fixed" and "curr_temp" in dict(fixed.named_buffers())
    assert list(fixed.state_dict().keys()) == ["curr_temp"]
    assert fixed.curr_temp.shape == (1,) and math.isclose(fixed.curr_temp.item(), 0.1, rel_tol=1e-6)
    learn = scp.SimpleVectorQuantizer("learnable=0.07")
    assert learn.temp_type == "learnable" and isinstance(learn.curr_temp, torch.nn.Parameter)
    assert list(learn.state_dict().keys()) == ["curr_temp"]
    sched = scp.SimpleVectorQuantizer("(2.0, 0.5, 0.9)")
    assert sched.temp_type == "scheduled"
    sched.set_num_updates(3)                                                     # my_vector_quantizer.py:58-62
    assert math.isclose(sched.curr_temp.item(), 2.0 * 0.9 ** 3, rel_tol=1e-6)
    sched.set_num_updates(1000)
    assert math.isclose(sched.curr_temp.item(), 0.5, rel_tol=1e-6)
    assert list(sched.state_dict().keys()) == []                                  # a python float in the reference
    fixed.set_num_updates(10)
    assert math.isclose(fixed.curr_temp.item(), 0.1, rel_tol=1e-6)
    with pytest.raises(NotImplementedError):
        scp.SimpleVectorQuantizer("fixed=0.1", use_gumbel=True)
    with pytest.raises(scp.ScpError):
        fixed(torch.zeros(1, 2, 8))                                              # CPU tensor
    gtp = scp.SimpleVectorQuantizer("fixed=0.1", groundTruthPerplexity=10.0)
    assert isinstance(gtp.perplexity_criteria, torch.nn.MSELoss)


def test_masked_array_validation():
    arr, n = _masked_array((0, 2, 3), 100)
    assert n == 3 and list(arr)[:3] == [0, 2, 3]
    arr, n = _masked_array((), 100)
    assert n == 0
    with pytest.raises(IndexError):
        _masked_array((0, 100), 100)
    with pytest.raises(scp.ScpError):
        _masked_array(tuple(range(9)), 100)


def test_contrastive_loss_contract():
    crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True)
    sd = crit.state_dict()
    assert set(sd.keys()) == {"temperature", "eye_mat", "neg_eye_mat", "eye_mat_fl"}   # losses.py:161-168
    assert sd["eye_mat"].shape == (256, 256) and sd["eye_mat"].dtype == torch.bool
    assert math.isclose(crit.temperature.item(), math.log(1 / 0.07), rel_tol=1e-6)
    assert math.isclose(crit.current_temperature, 1 / 0.07, rel_tol=1e-5)
    fixed = scp.MaskedContrastiveLoss(temperature=0.1)
    assert isinstance(fixed.temperature, float) and math.isclose(fixed.current_temperature, 10.0)
    assert set(fixed.state_dict().keys()) == {"eye_mat", "neg_eye_mat", "eye_mat_fl"}
    with pytest.raises(AssertionError):
        scp.MaskedContrastiveLoss(a2b=False, b2a=False)
    with pytest.raises(AssertionError):
        crit(torch.zeros(4, 8), torch.zeros(5, 8))                               # losses.py:199
    with pytest.raises(scp.ScpError):
        crit(torch.zeros(4, 64), torch.zeros(4, 64))
    # a checkpoint written by the reference class loads into the drop-in
    ref_like = {"temperature": torch.tensor(3.0), "eye_mat": torch.eye(256, dtype=torch.bool),
                "neg_eye_mat": ~torch.eye(256, dtype=torch.bool), "eye_mat_fl": torch.eye(256)}
    crit.load_state_dict(ref_like)
    assert crit.temperature.item() == 3.0


def test_pack_layout_roundtrip_on_host():
    n, D, n_feats, world = 6, 8, 3, 4
    assert kw_glue.pack_nbytes(n_feats, n, D) == n_feats * n * D * 4 + n * 8
    gen = torch.Generator().manual_seed(0)
    feats = [torch.randn(world * n, D, generator=gen) for _ in range(n_feats)]
    ids = torch.randint(0, 1000, (world * n,), generator=gen)
    rows = []
    for r in range(world):
        r0, r1 = kw_glue.shard_rows(n, r)
        parts = [f[r0:r1].contiguous().view(torch.uint8).reshape(-1) for f in feats]
        parts.append(ids[r0:r1].contiguous().view(torch.uint8).reshape(-1))
        rows.append(torch.cat(parts))
    gathered = torch.stack(rows)
    assert gathered.shape[1] == kw_glue.pack_nbytes(n_feats, n, D)
    out_feats, out_ids = kw_glue.unpack_gathered(gathered, n_feats, n, D)
    for a, b in zip(out_feats, feats):
        assert torch.equal(a, b)
    assert torch.equal(out_ids, ids)
    assert kw_glue.shard_rows(5, 3) == (15, 20)
    assert kw_glue.ddp_grad_scale(8) == 8.0
    # segment-major result of the peer all-gather (scp_p2p_allgather_segments): for every segment of the payload the world
    # copies follow each other in rank order -- restated here on the host -- and unpack_segments returns plain VIEWS
    fbytes = n * D * 4
    segs = [gathered[:, f * fbytes:(f + 1) * fbytes].reshape(-1) for f in range(n_feats)]
    segs.append(gathered[:, n_feats * fbytes:].reshape(-1))
    flat = torch.cat(segs)
    seg_feats, seg_ids = kw_glue.unpack_segments(flat, world, n_feats, n, D)
    for a, b in zip(seg_feats, feats):
        assert torch.equal(a, b) and a.data_ptr() >= flat.data_ptr() and a.is_contiguous()
    assert torch.equal(seg_ids, ids)
    assert kw_glue.gather_packed_segments(gathered[0], n_feats, n, D, kw_glue.LOCAL) is None   # world 1 / host tensors


def test_compute_loss_key_contract():
    calls = []

    def fake_criterion(feat_A, feat_B, index, **kw):
        calls.append((feat_A.shape, feat_B.shape, kw))
        return feat_A.sum() * 0 + 1.0

    feats = {"id": torch.arange(4), "image_feat": torch.ones(4, 8), "cascaded_audio_feat": torch.ones(4, 8),
             "parallel_audio_feat": torch.ones(4, 8)}
    out = scp.compute_loss(feats, fake_criterion, 1.5, 0.5)
    assert set(out) == {"loss", "c_cl_loss", "p_cl_loss"} and math.isclose(float(out["loss"]), 2.0)
    out = scp.compute_loss(feats, fake_criterion, 0.0, 1.0, local_rows=(0, 2))
    assert set(out) == {"loss", "p_cl_loss"} and calls[-1][2] == {"local_rows": (0, 2), "group": None}
    # quantity loss (kwClip.py:1030-1038): single process -> plain sum with the weight
    feats_q = dict(feats, cif_quantity_out=torch.tensor([1.0, 2.0]), cif_target_len=torch.tensor([2.0, 4.0]))
    out = scp.compute_loss(feats_q, fake_criterion, 1.0, 0.0, quantity_loss_weight=0.5,
                           quantity_loss_criteria=torch.nn.L1Loss())
    assert set(out) == {"loss", "c_cl_loss", "quantity_loss"} and math.isclose(float(out["loss"]), 1.0 + 0.5 * 1.5)
    with pytest.raises(AssertionError):
        scp.compute_loss({"id": torch.arange(4)}, fake_criterion, 1.0, 0.0)        # kwClip.py:1006-1010
