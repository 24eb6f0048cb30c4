"""Generate tests/golden/*.npz by running the REFERENCE's own modules on seeded inputs.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Each fixture stores the inputs and what the reference classes returned (outputs and autograd
gradients).  The fixtures are committed; the GPU box replays them without the reference.
Reference entry points exercised (paths relative to /root/reference):
  * avssl/module/weighted_sum.py            WeightedSumLayer.forward                      (S1)
  * avssl/module/speech_encoder_plus.py:518-622  FairseqSpeechEncoder_Hubert.forward with a stand-in upstream
                                            model: per-layer rescale, feat_len, weighted sum   (S1')
  * avssl/module/speechclip_c_modules/kw_bn.py  Kw_BatchNorm / Kw_BatchNorm_dynamic .forward   (N1)
  * avssl/model/kw_branches.py:158-197      GeneralBranch.get_keyword_cosine_score,
                                            GeneralBranch.vq_audio_features (identity projection)  (V1, V4)
  * avssl/module/speechclip_c_modules/my_vector_quantizer.py  SimpleVectorQuantizer.forward   (V3)
  * avssl/module/clip_official.py:222-279   ClipModel.encode_keywords around a stand-in text tower;
    avssl/util/data_utils.py:6-22           get_keypadding_mask                            (N3)
  * avssl/module/cif.py:97-311              CIF.forward / CIF.integrate_and_fire           (N4)
  * avssl/module/losses.py:129-245          MaskedContrastiveLoss                          (S3)
  * avssl/model/kwClip.py:999-1040          KWClip_GeneralTransformer.compute_loss         (C0)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as ref  # noqa: E402

SEED = 7122  # the reference's default seed, avssl/util/args.py:28


def _gen(seed_offset: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(SEED + seed_offset)
    return g


def _np(t):
    if torch.is_tensor(t):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def save(name: str, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: _np(v) for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


# ----------------------------------------------------------------------------------------------
def golden_wsum():
    ws = ref.load_leaf("avssl/module/weighted_sum.py", "ref_weighted_sum")
    cases = [
        # name, L, B, T, D, normalize, zero_weights
        ("wsum_base_plain", 13, 2, 7, 768, False, False),
        ("wsum_base_zero_w", 13, 2, 5, 768, False, True),
        ("wsum_large_norm", 25, 2, 5, 1024, True, False),
        ("wsum_small_norm", 4, 3, 9, 192, True, False),
    ]
    for i, (name, L, B, T, D, norm, zero_w) in enumerate(cases):
        g = _gen(100 + i)
        # HuBERT hands over (T,B,D)-storage tensors viewed as (B,T,D): speech_encoder_plus.py:596-599
        storage = [torch.randn(T, B, D, generator=g) * (1.0 + 0.3 * l) + 0.1 * l for l in range(L)]
        layers = [s.transpose(0, 1).requires_grad_(True) for s in storage]
        layer = ws.WeightedSumLayer(n_weights=L, normalize_features=norm)
        if not zero_w:
            with torch.no_grad():
                layer.weights.copy_(torch.randn(L, generator=g) * 0.5)
        y = layer(layers)
        gy = torch.randn(B, T, D, generator=g)
        grads = torch.autograd.grad(y, [layer.weights] + layers, grad_outputs=gy)
        save(name, layers_tbd=torch.stack(storage), weights=layer.weights, normalize=np.array(norm),
             y=y, grad_y=gy, grad_weights=grads[0], grad_layers=torch.stack([x for x in grads[1:]]))


# ----------------------------------------------------------------------------------------------
def golden_s1_tail():
    """Runs the reference's FairseqSpeechEncoder_Hubert.forward (speech_encoder_plus.py:518-622) around a stand-in
    upstream model that returns fixed layer_results: exercises the method1 / method2 rescale loop (:572-592), the
    feat_len arithmetic (:600-611) and the weighted-sum call (:619-622) exactly as shipped."""
    ref.import_avssl()
    import avssl.module.speech_encoder_plus as sep
    cases = [
        # name, L, B, T, D, normalize_hiddenstates, normalize_type, wav lengths
        ("s1tail_method1", 13, 3, 5, 384, True, "method1", [1600, 1120, 801]),   # 1120/320 = 3.5 -> 4 (half to even)
        ("s1tail_method2", 5, 3, 9, 256, True, "method2", [2880, 2400, 1761]),  # 2400/320 = 7.5 -> 8
        ("s1tail_method2_large", 25, 2, 3, 1024, True, "method2", [960, 800]),    # 800/320 = 2.5 -> 2
        ("s1tail_s3prl", 4, 3, 7, 192, True, "s3prl", [2240, 2239, 160]),
        ("s1tail_plain", 4, 2, 5, 64, False, "s3prl", [1600, 4000]),  # 4000/320 = 12.5 -> clamped to T
    ]
    for i, (name, L, B, T, D, norm, ntype, wav_lens) in enumerate(cases):
        g = _gen(500 + i)
        storage = [torch.randn(T, B, D, generator=g) * (1.0 + 0.3 * l) + 0.1 * l for l in range(L)]
        if name == "s1tail_method1":
            storage[3][2, 1] = 0.0  # an all-zero frame: x / (0 + 1e-8) = 0, gradient g / 1e-8
        layers = [s.transpose(0, 1).requires_grad_(True) for s in storage]

        class FakeUpstream(torch.nn.Module):
            def customHubertForward(self, wav, padding_mask=None, mask=None):
                return {"layer_results": list(layers)}

        enc = sep.FairseqSpeechEncoder_Hubert.__new__(sep.FairseqSpeechEncoder_Hubert)
        torch.nn.Module.__init__(enc)
        enc.encoder = FakeUpstream()
        enc.encoder_task = types.SimpleNamespace(cfg=types.SimpleNamespace(normalize=False))
        enc.trainable = True  # keeps the autograd graph through the rescale (else the upstream call is under no_grad)
        enc.normalize_hiddenstates = norm
        enc.normalize_type = ntype
        enc.downsample_rate = 320
        enc.max_audio_len = 102400
        enc.feat_select_idx = "weighted_sum"
        enc.weightedsum_layer = sep.WeightedSumLayer(n_weights=L, normalize_features=norm and ntype == "s3prl")
        with torch.no_grad():
            enc.weightedsum_layer.weights.copy_(torch.randn(L, generator=g) * 0.5)
        enc.eval()  # no random crop
        wav = [torch.randn(n, generator=g) for n in wav_lens]
        y, feat_len = enc(wav)
        gy = torch.randn(B, T, D, generator=g)
        grads = torch.autograd.grad(y, [enc.weightedsum_layer.weights] + layers, grad_outputs=gy)
        save(name, layers_tbd=torch.stack(storage), weights=enc.weightedsum_layer.weights,
             normalize_hiddenstates=np.array(norm), normalize_type=np.array(ntype), wav_len=np.array(wav_lens),
             y=y, feat_len=feat_len, grad_y=gy, grad_weights=grads[0],
             grad_layers=torch.stack([x for x in grads[1:]]))


# ----------------------------------------------------------------------------------------------
def golden_kwbn():
    """The reference's keyword batch-norm layers (kw_bn.py) in every configuration the branches can build: eachKw
    parallel (the shipped fixed-K recipes), eachKw with K separate layers, same, same with seq_lens, the dynamic layer
    of the "+" branches, plus an eval-mode case that normalises with (non-trivial) running statistics."""
    kb = ref.load_leaf("avssl/module/speechclip_c_modules/kw_bn.py", "ref_kw_bn")
    cases = [
        # name, B, K, D, type, parallel, training, seq_lens
        ("kwbn_eachkw_parallel", 6, 8, 64, "eachKw", True, True, None),
        ("kwbn_eachkw_layers", 5, 3, 32, "eachKw", False, True, None),
        ("kwbn_same", 4, 8, 64, "same", False, True, None),
        ("kwbn_same_seqlens", 4, 6, 32, "same", False, True, [6, 2, 4, 1]),
        ("kwbn_dynamic", 3, 11, 64, "dynamic", False, True, None),
        ("kwbn_eachkw_parallel_eval", 6, 8, 64, "eachKw", True, False, None),
    ]
    for i, (name, B, K, D, btype, parallel, training, seq_lens) in enumerate(cases):
        g = _gen(600 + i)
        table = torch.randn(300, D, generator=g) * 0.02 + 0.003 * torch.randn(1, D, generator=g)
        init_bias, init_scale = table.mean(0), table.std(0)   # kw_branches.py:99-100
        if btype == "dynamic":
            layer = kb.Kw_BatchNorm_dynamic(kw_dim=D, init_bias=init_bias, init_scale=init_scale, std_scale=1.0)
        else:
            layer = kb.Kw_BatchNorm(kw_num=K, kw_dim=D, batchnorm_type=btype, init_bias=init_bias,
                                    init_scale=init_scale, std_scale=1.0, learnable=True, parallel=parallel)
        bns = list(layer.bn_layers) if hasattr(layer, "bn_layers") else [layer.bn_layer]
        for bn in bns:  # non-trivial running statistics
            with torch.no_grad():
                bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.1)
                bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
        rm0 = torch.stack([bn.running_mean.clone() for bn in bns])
        rv0 = torch.stack([bn.running_var.clone() for bn in bns])
        layer.train(training)
        x = (torch.randn(B, K, D, generator=g) * 0.7 + 0.3 * torch.randn(1, 1, D, generator=g)).requires_grad_(True)
        x_in = x.clone()  # the seq_lens branch writes into its input (kw_bn.py:157)
        if seq_lens is not None:
            y = layer(x_in, torch.tensor(seq_lens))
        else:
            y = layer(x_in)
        gy = torch.randn(B, K, D, generator=g)
        params = [p for bn in bns for p in (bn.weight, bn.bias)]
        grads = torch.autograd.grad(y, [x] + params, grad_outputs=gy)
        save(name, x=x, batchnorm_type=np.array(btype), parallel=np.array(parallel), training=np.array(training),
             seq_lens=np.array(seq_lens if seq_lens is not None else [], dtype=np.int64),
             weight=torch.stack([bn.weight for bn in bns]), bias=torch.stack([bn.bias for bn in bns]),
             running_mean_in=rm0, running_var_in=rv0,
             running_mean_out=torch.stack([bn.running_mean for bn in bns]),
             running_var_out=torch.stack([bn.running_var for bn in bns]),
             num_batches_tracked=torch.stack([bn.num_batches_tracked for bn in bns]),
             y=y, grad_y=gy, grad_x=grads[0], grad_weight=torch.stack(list(grads[1::2])),
             grad_bias=torch.stack(list(grads[2::2])))


# ----------------------------------------------------------------------------------------------
def golden_splice():
    """ClipModel.encode_keywords (clip_official.py:222-279) run unbound on a duck-typed CLIP whose "transformer" records
    its input: pins the spliced (B,77,D) tensor, the EOT gather and the gradient that reaches the keywords."""
    ref.import_avssl()
    import avssl.module.clip_official as co
    import avssl.util.data_utils as du
    cases = [
        # name, B, Kmax, D, V, keyword_num (int or per-utterance list), reduced vocabulary
        ("splice_fixed8", 3, 8, 64, 120, 8, False),
        ("splice_dynamic", 4, 12, 64, 120, [12, 3, 7, 1], False),
        ("splice_dynamic_reduced", 3, 6, 32, 90, [2, 6, 4], True),
    ]
    for i, (name, B, Kmax, D, V, num, reduced) in enumerate(cases):
        g = _gen(700 + i)
        L = 77
        emb = torch.nn.Embedding(V, D)
        with torch.no_grad():
            emb.weight.copy_(torch.randn(V, D, generator=g) * 0.02)
        emb.weight.requires_grad_(False)
        captured = {}

        class Tower(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.mix = torch.nn.Linear(D, D)

            def forward(self, x):  # (L, N, D)
                captured["x"] = x.permute(1, 0, 2)
                return torch.tanh(self.mix(x))

        torch.manual_seed(SEED + 700 + i)
        model = types.SimpleNamespace(token_embedding=emb, positional_embedding=torch.randn(L, D, generator=g) * 0.01,
                                      transformer=Tower(), ln_final=torch.nn.LayerNorm(D),
                                      text_projection=torch.randn(D, 16, generator=g) * 0.1)
        sot, eot = V - 2, V - 1
        clip = types.SimpleNamespace(model=model, device=torch.device("cpu"),
                                     tokenizer=types.SimpleNamespace(encoder={"<|startoftext|>": sot, "<|endoftext|>": eot}),
                                     selected_text_emb_ids=None)
        if reduced:  # ids come from the reduced-vocabulary attributes instead of the tokenizer (:246-247)
            clip.selected_text_emb_ids = np.arange(V)
            clip.startOfTxt_reduced, clip.endOfTxt_reduced = 5, 7
            sot, eot = 5, 7
        kw = (torch.randn(B, Kmax, D, generator=g) * 0.02).requires_grad_(True)
        keyword_num = torch.tensor(num) if isinstance(num, list) else num
        out = co.ClipModel.encode_keywords(clip, kw, keyword_num)
        g_out = torch.randn(out.shape, generator=g)
        (g_kw,) = torch.autograd.grad(out, [kw], grad_outputs=g_out)
        lens = torch.tensor(num) if isinstance(num, list) else torch.full((B,), num)
        mask = du.get_keypadding_mask(Kmax, lens)
        save(name, keywords=kw, keyword_num=np.array(num), table=emb.weight, pos_emb=model.positional_embedding,
             sot=np.array(sot), eot=np.array(eot), x=captured["x"], mix_w=model.transformer.mix.weight,
             mix_b=model.transformer.mix.bias, text_projection=model.text_projection, out=out, grad_out=g_out,
             grad_keywords=g_kw, keypadding_mask=mask)


# ----------------------------------------------------------------------------------------------
def golden_cif():
    """The reference's CIF down-sampler (cif.py): integrate_and_fire in training mode (target lengths, scaling, tail
    ignored), with several fires per source, in inference mode (tail handling with and without an extra fire), without
    tail handling, and one full CIF.forward (conv weight generator + scaling + integrate)."""
    cif_mod = ref.load_leaf("avssl/module/cif.py", "ref_cif")
    cases = [
        # name, B, S, C, mode, alpha scale
        ("cif_train_scaled", 4, 40, 64, "train", 0.3),
        ("cif_train_multifire", 3, 6, 32, "train_multifire", 0.5),
        ("cif_infer_tail", 6, 37, 64, "infer", 0.35),
        ("cif_infer_notail", 3, 25, 32, "notail", 0.4),
    ]
    for i, (name, B, S, C, mode, a_scale) in enumerate(cases):
        g = _gen(800 + i)
        x = torch.randn(B, S, C, generator=g).requires_grad_(True)
        raw = (torch.rand(B, S, generator=g) * 2 * a_scale)
        lens = torch.randint(S // 2, S + 1, (B,), generator=g)
        lens[0] = S
        pad = torch.arange(S)[None, :] >= lens[:, None]
        raw = raw.masked_fill(pad, 0.0).requires_grad_(True)
        layer = cif_mod.CIF(cif_threshold=1.0, cif_output_dim=C, encoder_embed_dim=C,
                            apply_tail_handling=(mode != "notail"))
        target = None
        alpha = raw
        if mode.startswith("train"):
            target = torch.tensor([12, 5, 9, 3][:B]) if mode == "train" else torch.tensor([10, 13, 7])
            alpha = raw * ((1.0 * target.type_as(raw) + 1e-5) / raw.sum(1)).unsqueeze(1)   # the scaling of cif.py:126-129
        out = layer.integrate_and_fire(x, alpha, target_lengths=target)
        feats = out["dsample_feats"]
        gy = torch.randn(feats.shape, generator=g)
        if mode == "infer":
            # the reference updates fire_mask in place after autograd saved it (cif.py:281-283): its inference tail path
            # cannot be back-propagated -- forward values only
            gx, ga = torch.zeros_like(x), torch.zeros_like(raw)
        else:
            gx, ga = torch.autograd.grad(feats, [x, raw], grad_outputs=gy)
        save(name, x=x, alpha_raw=raw, mode=np.array(mode), target_len=(target if target is not None else np.array([], dtype=np.int64)),
             apply_tail_handling=np.array(mode != "notail"), alpha=alpha, feats=feats, feat_len=out["dsample_feats_length"],
             pad_mask=out["dsample_feats_pad_mask"], fired_marks=out["fired_marks"], grad_feats=gy, grad_x=gx,
             grad_alpha_raw=ga)
    # full module forward: conv weight generator -> clip / mask -> scaling -> integrate (training-style call, eval() so
    # that the two nn.Dropout layers are inert)
    g = _gen(850)
    B, S, C = 3, 30, 32
    torch.manual_seed(SEED + 850)
    layer = cif_mod.CIF(cif_threshold=1.0, cif_output_dim=C, encoder_embed_dim=C, conv_cif_width=3, scaling_step=100).eval()
    x = torch.randn(B, S, C, generator=g).requires_grad_(True)
    lens = torch.tensor([30, 21, 26])
    pad = torch.arange(S)[None, :] >= lens[:, None]
    target = torch.tensor([4, 2, 3])
    res = layer({"audio_feat": x, "audio_feat_pad_mask": pad, "global_step": 0}, target)
    gy = torch.randn(res["dsample_feats"].shape, generator=g)
    params = list(layer.parameters())
    grads = torch.autograd.grad(res["dsample_feats"].mul(gy).sum() + res["quantity_out"].sum(), [x] + params)
    save("cif_forward_train", x=x, pad_mask_in=pad, target_len=target, feats=res["dsample_feats"],
         feat_len=res["dsample_feats_length"], quantity_out=res["quantity_out"], orig_alpha=res["orig_alpha"],
         alpha=res["alpha"], original_length=res["original_length"], grad_feats=gy, grad_x=grads[0],
         **{f"param_{n.replace('.', '_')}": p for n, p in layer.named_parameters()},
         **{f"grad_{n.replace('.', '_')}": gr for (n, _), gr in zip(layer.named_parameters(), grads[1:])})


# ----------------------------------------------------------------------------------------------
def _fake_branch(kb, vq_mod, table: torch.Tensor, temp_spec: str, training: bool, hard: bool = True):
    """A GeneralBranch whose projection is the identity, so that vq_audio_features (kw_branches.py:181-197)
    runs V1 + V3 + V4 of the reference on the given keyword vectors."""
    branch = kb.GeneralBranch.__new__(kb.GeneralBranch)
    torch.nn.Module.__init__(branch)
    branch.text_dim = table.shape[1]
    emb = torch.nn.Embedding(table.shape[0], table.shape[1])
    with torch.no_grad():
        emb.weight.copy_(table)
    emb.weight.requires_grad_(False)
    branch.clip = types.SimpleNamespace(model=types.SimpleNamespace(token_embedding=emb))
    branch.linear_proj = torch.nn.Identity()
    branch.vector_quantizer = vq_mod.SimpleVectorQuantizer(temp=temp_spec, hard=hard)
    branch.train(training)
    return branch


def golden_vq():
    ref.import_avssl()
    import avssl.model.kw_branches as kb
    vq_mod = ref.load_leaf("avssl/module/speechclip_c_modules/my_vector_quantizer.py", "ref_vq")
    cases = [
        # name, B, K, V, D, temp_spec, training, duplicate rows in the table
        ("vq_train_fixed", 3, 4, 512, 64, "fixed=0.1", True, False),
        ("vq_eval_fixed", 3, 4, 512, 64, "fixed=0.1", False, False),
        ("vq_train_learnable", 2, 8, 1024, 128, "learnable=0.07", True, False),
        ("vq_train_ties", 2, 3, 260, 64, "fixed=0.1", True, True),
        # hard=False (my_vector_quantizer.py:130-136 without the straight-through term): subword_prob = softmax(x / tau)
        ("vq_train_soft", 3, 4, 512, 64, "fixed=0.1", True, False, False),
    ]
    for i, case in enumerate(cases):
        name, B, K, V, D, temp_spec, training, dup = case[:8]
        hard = case[8] if len(case) > 8 else True
        g = _gen(200 + i)
        table = torch.randn(V, D, generator=g) * 0.02 + 0.003 * torch.randn(1, D, generator=g)
        if dup:
            # exact duplicates -> exact ties: "first max wins" must hold (my_vector_quantizer.py:82)
            table[17] = table[200]
            table[5] = table[200]
        # keywords mimic the batch-norm output initialised to table mean/std (kw_branches.py:99-100)
        kw = torch.randn(B, K, D, generator=g) * table.std(0) + table.mean(0)
        if dup:
            kw[0, 0] = table[200] * 3.0
            kw[1, 2] = table[2] * 2.0  # best raw match is a masked column (2): must not be selected
        kw.requires_grad_(True)
        branch = _fake_branch(kb, vq_mod, table, temp_spec, training, hard)
        cos = branch.get_keyword_cosine_score(kw.detach())
        vq_results, kw_out = branch.vq_audio_features(kw)
        arrays = dict(keywords_in=kw, table=table, training=np.array(training), temp_spec=np.array(temp_spec),
                      hard=np.array(hard),
                      cos=cos, keywords_out=kw_out, subword_prob=vq_results["subword_prob"],
                      targets=vq_results["targets"], code_perplexity=vq_results["code_perplexity"],
                      prob_perplexity=vq_results["prob_perplexity"], ent_per_t=vq_results["ent_per_t"],
                      temp=np.array(vq_results["temp"]), diversity_loss=vq_results["diversity_loss"],
                      num_vars=np.array(vq_results["num_vars"]))
        if training:
            g_out = torch.randn(B, K, D, generator=g)
            params = [kw]
            if temp_spec.startswith("learnable"):
                params.append(branch.vector_quantizer.curr_temp)
            grads = torch.autograd.grad(kw_out, params, grad_outputs=g_out)
            arrays.update(grad_keywords_out=g_out, grad_keywords_in=grads[0])
            if len(grads) > 1:
                arrays.update(grad_temp=grads[1])
        save(name, **arrays)


# ----------------------------------------------------------------------------------------------
def golden_nce():
    losses = ref.load_leaf("avssl/module/losses.py", "ref_losses")
    cases = [
        # name, N, D, ctor kwargs, ids kind
        ("nce_n8_trainable", 8, 64, dict(temperature=0.07, temperature_trainable=True), "unique"),
        ("nce_n48_dupids", 48, 64, dict(temperature=0.07, temperature_trainable=True), "dup"),
        ("nce_n48_fixed_margin", 48, 128, dict(temperature=0.1, temperature_trainable=False, margin=0.2), "dup"),
        ("nce_n40_dcl", 40, 64, dict(temperature=0.07, temperature_trainable=True, dcl=True), "dup"),
        ("nce_n40_a2b_only", 40, 64, dict(temperature=0.07, temperature_trainable=True, b2a=False), "dup"),
        ("nce_n40_b2a_noindex", 40, 64, dict(temperature=0.07, temperature_trainable=True, a2b=False), "none"),
        ("nce_n300_big", 300, 64, dict(temperature=0.07, temperature_trainable=True), "dup"),
    ]
    for i, (name, N, D, kw, ids_kind) in enumerate(cases):
        g = _gen(300 + i)
        a = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1).requires_grad_(True)
        b = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + 0.5 * a.detach(), dim=-1).requires_grad_(True)
        if ids_kind == "unique":
            ids = torch.arange(N)
        elif ids_kind == "dup":
            ids = torch.randint(0, max(N // 5, 1), (N,), generator=g)  # Flickr: 5 captions per image
        else:
            ids = None
        if N > losses.MAX_EYE:
            losses.MAX_EYE = N  # reference limit losses.py:126 (IndexError for N > 256)
        crit = losses.MaskedContrastiveLoss(**kw)
        loss = crit(a, b, ids)
        params = [a, b] + ([crit.temperature] if kw.get("temperature_trainable") else [])
        grads = torch.autograd.grad(loss, params)
        arrays = dict(feat_a=a, feat_b=b, ids=ids if ids is not None else np.array([-1]),
                      has_ids=np.array(ids is not None), loss=loss, grad_a=grads[0], grad_b=grads[1],
                      temperature=np.array(kw["temperature"]), trainable=np.array(kw.get("temperature_trainable", False)),
                      margin=np.array(kw.get("margin", 0.0)), dcl=np.array(kw.get("dcl", False)),
                      a2b=np.array(kw.get("a2b", True)), b2a=np.array(kw.get("b2a", True)),
                      current_temperature=np.array(crit.current_temperature))
        if len(grads) > 2:
            arrays["grad_temperature"] = grads[2]
        save(name, **arrays)


# ----------------------------------------------------------------------------------------------
def golden_hybrid_loss():
    ref.import_avssl()
    import avssl.model.kwClip as kc
    losses = ref.load_leaf("avssl/module/losses.py", "ref_losses2")
    g = _gen(400)
    N, D = 32, 64
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    ca = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + img, dim=-1).requires_grad_(True)
    pa = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + 2 * img, dim=-1).requires_grad_(True)
    ids = torch.randint(0, 8, (N,), generator=g)
    qo = torch.rand(N, generator=g) * 10
    ql = torch.randint(4, 12, (N,), generator=g).float()
    fake = types.SimpleNamespace()
    fake.config = types.SimpleNamespace(model_settings=types.SimpleNamespace(
        cascaded_objective_weight=1.5, parallel_objective_weight=0.5))
    fake.criterion = losses.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True)
    fake.quantity_loss_criteria = torch.nn.L1Loss()
    fake.quantity_loss_weight = 0.25
    out = kc.KWClip_GeneralTransformer.compute_loss(
        fake, {"id": ids, "image_feat": img, "cascaded_audio_feat": ca, "parallel_audio_feat": pa,
               "cif_quantity_out": qo, "cif_target_len": ql})
    grads = torch.autograd.grad(out["loss"], [ca, pa, fake.criterion.temperature])
    save("hybrid_loss", image_feat=img, cascaded_audio_feat=ca, parallel_audio_feat=pa, ids=ids,
         cif_quantity_out=qo, cif_target_len=ql, cascaded_weight=np.array(1.5), parallel_weight=np.array(0.5),
         quantity_loss_weight=np.array(0.25), temperature=np.array(0.07),
         loss=out["loss"], c_cl_loss=out["c_cl_loss"], p_cl_loss=out["p_cl_loss"], quantity_loss=out["quantity_loss"],
         grad_cascaded=grads[0], grad_parallel=grads[1], grad_temperature=grads[2])

# ----------------------------------------------------------------------------------------------
def golden_chain():
    """The cascaded branch from the CLS keywords to the loss, with the reference's own glue in between
    (kw_branches.py:143-156, :181-197, :390-395 / :744-750; clip_official.py:222-279; kwClip.py:905-907, :1015-1028):

        audio_feat (B,K,Da) -> linear_proj -> Kw_BatchNorm[_dynamic] -> cosine -> VQ -> subword_prob @ E
                            -> ClipModel.encode_keywords (stand-in text tower) -> x / ||x|| -> MaskedContrastiveLoss

    Pins the COMPOSITION of N1 + V1 + V3 + V4 + N3 + N0 + S3 and the gradient that flows back through all of them to the
    projection, the batch-norm parameters, the text tower and the loss temperature."""
    ref.import_avssl()
    import avssl.model.kw_branches as kb
    import avssl.module.clip_official as co
    vq_mod = ref.load_leaf("avssl/module/speechclip_c_modules/my_vector_quantizer.py", "ref_vq_chain")
    bn_mod = ref.load_leaf("avssl/module/speechclip_c_modules/kw_bn.py", "ref_kw_bn_chain")
    losses = ref.load_leaf("avssl/module/losses.py", "ref_losses_chain")
    cases = [
        # name, B, K, Da, D, V, dynamic keyword counts (None = fixed K for every utterance)
        ("chain_cascaded_fixed8", 16, 8, 96, 64, 600, None),
        ("chain_cascaded_dynamic", 12, 10, 96, 64, 600, [10, 3, 7, 1, 5, 10, 2, 8, 4, 6, 9, 3]),
    ]
    for i, (name, B, K, Da, D, V, lens) in enumerate(cases):
        g = _gen(800 + i)
        torch.manual_seed(SEED + 800 + i)
        L = 77
        table = torch.randn(V, D, generator=g) * 0.02 + 0.003 * torch.randn(1, D, generator=g)
        emb = torch.nn.Embedding(V, D)
        with torch.no_grad():
            emb.weight.copy_(table)
        emb.weight.requires_grad_(False)

        class Tower(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.mix = torch.nn.Linear(D, D)

            def forward(self, x):  # (L, N, D)
                return torch.tanh(self.mix(x))

        model = types.SimpleNamespace(token_embedding=emb, positional_embedding=torch.randn(L, D, generator=g) * 0.01,
                                      transformer=Tower(), ln_final=torch.nn.LayerNorm(D),
                                      text_projection=torch.randn(D, 64, generator=g) * 0.1)
        sot, eot = V - 2, V - 1
        clip = types.SimpleNamespace(model=model, device=torch.device("cpu"), selected_text_emb_ids=None,
                                     tokenizer=types.SimpleNamespace(encoder={"<|startoftext|>": sot,
                                                                              "<|endoftext|>": eot}))
        clip.encode_keywords = types.MethodType(co.ClipModel.encode_keywords, clip)

        branch = kb.GeneralBranch.__new__(kb.GeneralBranch)
        torch.nn.Module.__init__(branch)
        branch.text_dim = D
        branch.clip = clip
        branch.linear_proj = torch.nn.Linear(Da, D)
        init_bias, init_scale = table.mean(0), table.std(0)                       # kw_branches.py:99-100
        if lens is None:
            branch.bn_layer = bn_mod.Kw_BatchNorm(kw_num=K, kw_dim=D, batchnorm_type="eachKw", init_bias=init_bias,
                                                  init_scale=init_scale, std_scale=1, learnable=True, parallel=True)
        else:
            branch.bn_layer = bn_mod.Kw_BatchNorm_dynamic(kw_dim=D, init_bias=init_bias, init_scale=init_scale,
                                                          std_scale=1, learnable=True)
        branch.vector_quantizer = vq_mod.SimpleVectorQuantizer(temp="fixed=0.1")
        branch.train(True)
        bn_state_in = {k: v.clone() for k, v in branch.bn_layer.state_dict().items()}
        criterion = losses.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True)

        audio_feat = torch.randn(B, K, Da, generator=g).requires_grad_(True)
        image_feat = torch.randn(B, 64, generator=g)
        image_feat = image_feat / image_feat.norm(dim=-1, keepdim=True)           # kwClip.py:857
        ids = torch.randint(0, B // 2, (B,), generator=g)

        # -- the branch forward from the CLS keywords on (kw_branches.py:390-395 / :744-750)
        vq_results, keywords = branch.vq_audio_features(audio_feat)
        keyword_num = K if lens is None else torch.tensor(lens)
        cascaded = clip.encode_keywords(keywords, keyword_num)
        cascaded_n = cascaded / cascaded.norm(dim=-1, keepdim=True)               # kwClip.py:905-907
        loss = criterion(feat_A=cascaded_n.float(), feat_B=image_feat.float(), index=ids)  # kwClip.py:1021-1025

        bn_state_out = {k: v.clone() for k, v in branch.bn_layer.state_dict().items()}

        # the chain is only a meaningful pin if no arg-max sits on a knife edge: check the fp64 top-2 gap
        # (this second training-mode call moves the running statistics again: they were snapshotted above)
        with torch.no_grad():
            bn_out = branch.project_feats_to_CLIPspace(audio_feat).double()
            cos = torch.nn.functional.normalize(bn_out, dim=-1) @ torch.nn.functional.normalize(table.double(), dim=-1).t()
            cos[..., [0, 2, 3]] = -float("inf")
            top2 = cos.topk(2, dim=-1).values
            assert (top2[..., 0] - top2[..., 1]).min() > 1e-5, "regenerate with another seed: near-tie in the fixture"

        params = {"audio_feat": audio_feat, "proj_weight": branch.linear_proj.weight, "proj_bias": branch.linear_proj.bias,
                  "bn_weight": branch.bn_layer.bn_layer.weight, "bn_bias": branch.bn_layer.bn_layer.bias,
                  "mix_w": model.transformer.mix.weight, "mix_b": model.transformer.mix.bias,
                  "ln_weight": model.ln_final.weight, "ln_bias": model.ln_final.bias,
                  "temperature": criterion.temperature}
        grads = torch.autograd.grad(loss, list(params.values()))
        save(name, audio_feat=audio_feat, image_feat=image_feat, ids=ids,
             keyword_num=np.array(lens if lens is not None else K),
             table=table, pos_emb=model.positional_embedding, sot=np.array(sot), eot=np.array(eot),
             proj_weight=branch.linear_proj.weight, proj_bias=branch.linear_proj.bias,
             mix_w=model.transformer.mix.weight, mix_b=model.transformer.mix.bias,
             text_projection=model.text_projection,
             **{f"bn_in__{k.replace('.', '__')}": v for k, v in bn_state_in.items()},
             **{f"bn_out__{k.replace('.', '__')}": v for k, v in bn_state_out.items()},
             targets=vq_results["targets"], keywords=keywords, code_perplexity=vq_results["code_perplexity"],
             prob_perplexity=vq_results["prob_perplexity"], ent_per_t=vq_results["ent_per_t"],
             cascaded_audio_feat=cascaded, loss=loss, temperature=np.array(0.07),
             **{f"grad_{k}": gr for k, gr in zip(params, grads)})


if __name__ == "__main__":
    assert ref.reference_available(), "needs /root/reference"
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    groups = {"wsum": golden_wsum, "s1tail": golden_s1_tail, "kwbn": golden_kwbn, "splice": golden_splice,
              "cif": golden_cif, "vq": golden_vq, "nce": golden_nce, "hybrid_loss": golden_hybrid_loss,
              "chain": golden_chain}
    for key in (sys.argv[1:] or list(groups)):   # `make_golden.py chain` regenerates one group only
        groups[key]()
