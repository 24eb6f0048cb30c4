"""Register the B200 kernels in the reference's own plugin points.

The reference resolves the three hot-path classes by name on Python modules:
  * ``getattr(losses, config.cl_loss.type)``                      avssl/model/kwClip.py:84
  * ``getattr(vector_quantizers, vq.type)``                       avssl/model/kw_branches.py:75-91
  * ``WeightedSumLayer(...)`` imported by name                    avssl/module/speech_encoder_plus.py:24, :218-220, :472-476
  * the cosine + lookup glue is a method of ``GeneralBranch``     avssl/model/kw_branches.py:158-197
  * ``Kw_BatchNorm`` / ``Kw_BatchNorm_dynamic`` imported by name  avssl/model/kw_branches.py:19, :95, :629
  * ``CIF`` imported by name                                      avssl/model/kw_branches.py:17, :617
and, so that the glue between those modules is on the path too (model/kwclip_glue.py):
  * ``KWClip_GeneralTransformer.compute_loss`` / ``.forward``     avssl/model/kwClip.py:999-1040, :839-960  (N0, G0, C0)
  * ``FairseqSpeechEncoder_Hubert.forward`` / ``S3prlSpeechEncoderPlus.forward``  speech_encoder_plus.py:520-640, :240-311 (S1')
  * ``ClipModel.__init__``                                        avssl/module/clip_official.py:30-108  (N2)

``install()`` must run after ``import avssl`` and before the model is constructed; ``uninstall()`` restores the reference.
"""
from __future__ import annotations

import importlib
import sys

_MISSING = object()
_UNDO: list = []          # (owner, attribute, value before the first install()) in patch order
_UNDO_KEYS: set = set()


def _remember(owner, attr: str) -> None:
    key = (id(owner), attr)
    if key in _UNDO_KEYS:
        return
    _UNDO_KEYS.add(key)
    _UNDO.append((owner, attr, owner.__dict__.get(attr, _MISSING) if isinstance(owner, type) else getattr(owner, attr, _MISSING)))


def uninstall() -> int:
    """Put back everything ``install()`` replaced (classes, methods, wrappers); returns the number of restored names."""
    n = 0
    while _UNDO:
        owner, attr, old = _UNDO.pop()
        if old is _MISSING:
            if attr in getattr(owner, "__dict__", {}):
                delattr(owner, attr)
        else:
            setattr(owner, attr, old)
        n += 1
    _UNDO_KEYS.clear()
    return n


def install(avssl_package: str = "avssl", strict: bool = False, fuse_forward: bool = True) -> dict:
    """Swap the reference classes for the CUDA-backed ones.  Returns {dotted name: replaced?}.

    ``fuse_forward=False`` keeps the reference's own ``KWClip_GeneralTransformer.forward`` (its three feature
    normalisations then run as torch ops and the pack kernel re-normalises unit rows, which is the identity)."""
    from .module.losses import MaskedContrastiveLoss
    from .module.vector_quantizers import SimpleVectorQuantizer, fused_vq_audio_features
    from .module.weighted_sum import WeightedSumLayer
    from .module.kw_bn import Kw_BatchNorm, Kw_BatchNorm_dynamic
    from .module.cif import CIF

    done = {}

    def patch(mod_name: str, attr: str, value) -> None:
        key = f"{mod_name}.{attr}"
        try:
            mod = sys.modules.get(mod_name) or importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            done[key] = False
            return
        _remember(mod, attr)
        setattr(mod, attr, value)
        done[key] = True

    p = avssl_package
    patch(f"{p}.module.losses", "MaskedContrastiveLoss", MaskedContrastiveLoss)
    patch(f"{p}.module", "MaskedContrastiveLoss", MaskedContrastiveLoss)
    patch(f"{p}.module.weighted_sum", "WeightedSumLayer", WeightedSumLayer)
    patch(f"{p}.module.speech_encoder_plus", "WeightedSumLayer", WeightedSumLayer)
    patch(f"{p}.module", "WeightedSumLayer", WeightedSumLayer)
    patch(f"{p}.module.speechclip_c_modules.my_vector_quantizer", "SimpleVectorQuantizer", SimpleVectorQuantizer)
    patch(f"{p}.module.speechclip_c_modules.vector_quantizers", "SimpleVectorQuantizer", SimpleVectorQuantizer)
    for name, cls in (("Kw_BatchNorm", Kw_BatchNorm), ("Kw_BatchNorm_dynamic", Kw_BatchNorm_dynamic)):
        patch(f"{p}.module.speechclip_c_modules.kw_bn", name, cls)
        patch(f"{p}.model.kw_branches", name, cls)  # imported by name there (kw_branches.py:19)
    patch(f"{p}.module.cif", "CIF", CIF)                 # N4: imported by name in kw_branches.py
    patch(f"{p}.model.kw_branches", "CIF", CIF)
    # N3: the text-transformer input splice replaces the body of ClipModel.encode_keywords; get_keypadding_mask by name
    try:
        from .module.clip_glue import encode_keywords, get_keypadding_mask
        co = sys.modules.get(f"{p}.module.clip_official") or importlib.import_module(f"{p}.module.clip_official")
        _remember(co.ClipModel, "encode_keywords")
        co.ClipModel.encode_keywords = encode_keywords
        done[f"{p}.module.clip_official.ClipModel.encode_keywords"] = True
        patch(f"{p}.util.data_utils", "get_keypadding_mask", get_keypadding_mask)
        patch(f"{p}.model.kw_branches", "get_keypadding_mask", get_keypadding_mask)
    except Exception:
        if strict:
            raise
        done[f"{p}.module.clip_official.ClipModel.encode_keywords"] = False
    # the fused cosine + quantise + lookup replaces the method body on the shared base class of all branches
    try:
        kb = sys.modules.get(f"{p}.model.kw_branches") or importlib.import_module(f"{p}.model.kw_branches")
        _remember(kb.GeneralBranch, "vq_audio_features")
        kb.GeneralBranch.vq_audio_features = lambda self, audio_feat: fused_vq_audio_features(self, audio_feat)
        done[f"{p}.model.kw_branches.GeneralBranch.vq_audio_features"] = True
    except Exception:
        if strict:
            raise
        done[f"{p}.model.kw_branches.GeneralBranch.vq_audio_features"] = False

    # ---- the glue around the modules: loss (N0 + G0 + C0), upstream tail (S1'), vocabulary reduction (N2)
    from .model import kwclip_glue

    def patch_method(mod_name: str, cls_name: str, attr: str, make) -> None:
        key = f"{mod_name}.{cls_name}.{attr}"
        try:
            mod = sys.modules.get(mod_name) or importlib.import_module(mod_name)
            cls = getattr(mod, cls_name)
            current = cls.__dict__[attr] if attr in cls.__dict__ else getattr(cls, attr)
            if getattr(current, "_scp_installed", False):  # install() twice: do not wrap the wrapper
                done[key] = True
                return
            new = make(current)
            new._scp_installed = True
            new._scp_original = current
            _remember(cls, attr)
            setattr(cls, attr, new)
            done[key] = True
        except Exception:
            if strict:
                raise
            done[key] = False

    patch_method(f"{p}.model.kwClip", "KWClip_GeneralTransformer", "compute_loss",
                 lambda cur: (lambda self, inputDict: kwclip_glue.kwclip_compute_loss(self, inputDict)))
    if fuse_forward:
        patch_method(f"{p}.model.kwClip", "KWClip_GeneralTransformer", "forward",
                     lambda cur: (lambda self, batch: kwclip_glue.kwclip_forward(self, batch)))
    for cls_name in ("FairseqSpeechEncoder_Hubert", "S3prlSpeechEncoderPlus"):
        patch_method(f"{p}.module.speech_encoder_plus", cls_name, "forward", kwclip_glue.fused_upstream_forward)
    patch_method(f"{p}.module.clip_official", "ClipModel", "__init__", kwclip_glue.clipmodel_init)
    return done
