"""speechclip_plus_b200 -- B200-native (sm_100a) implementation of the SpeechCLIP+ data-parallel training hot path.

Public surface (mirrors the reference's ``avssl.module`` names for this path):

    WeightedSumLayer           avssl/module/weighted_sum.py
    SimpleVectorQuantizer      avssl/module/speechclip_c_modules/my_vector_quantizer.py
    fused_vq_audio_features    body of GeneralBranch.vq_audio_features, avssl/model/kw_branches.py:181-197
    MaskedContrastiveLoss      avssl/module/losses.py
    Kw_BatchNorm(_dynamic)     avssl/module/speechclip_c_modules/kw_bn.py          (keyword batch-norm before the VQ)
    CIF                        avssl/module/cif.py                                  (integrate-and-fire down-sampler)
    fuse_upstream_features     caller tail of the HuBERT wrapper, avssl/module/speech_encoder_plus.py:572-622
    gather_loss_feats / compute_loss     gather point + loss of avssl/model/kwClip.py:149-193, :999-1040
    PackedAdam                 packed gradient all-reduce + fused Adam of the path's own parameters (kwClip.py:636-668)
    install()                  registers the above in the reference's plugin namespaces

Everything computes in libscp_b200.so (hand-written CUDA for sm_100a, C ABI in include/scp_b200.h); there is no
CPU or eager-PyTorch fallback -- a missing library or a non-CUDA tensor raises ``ScpError``.
"""
from ._lib import ScpError, load as load_library, num_launches  # noqa: F401
from .module.weighted_sum import WeightedSumLayer  # noqa: F401
from .module.vector_quantizers import SimpleVectorQuantizer, TokenTableCache, fused_vq_audio_features  # noqa: F401
from .module.losses import MaskedContrastiveLoss  # noqa: F401
from .module.kw_bn import Kw_BatchNorm, Kw_BatchNorm_dynamic  # noqa: F401
from .module.speech_encoder_plus import fuse_upstream_features, upstream_feat_len  # noqa: F401
from .module.cif import CIF  # noqa: F401
from .model.kw_glue import compute_loss, gather_loss_feats, ddp_grad_scale  # noqa: F401
from .model.packed_optim import PackedAdam  # noqa: F401
from .install import install, uninstall  # noqa: F401

__version__ = "0.2.0"
