"""CPU: the C-ABI library loads and exports every symbol include/scp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scp_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|size_t|int64_t|const char\*)\s+(scp_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        decls[m.group(1)] = n
    return decls


def test_header_declares_the_expected_entry_points():
    decls = _declared_functions()
    for name in ["scp_wsum_fwd", "scp_wsum_bwd", "scp_vq_prepare_table", "scp_vq_fwd", "scp_vq_bwd", "scp_vq_dense_fwd",
                 "scp_vq_dense_bwd", "scp_l2norm_pack", "scp_l2norm_bwd", "scp_nce_fwd", "scp_nce_bwd", "scp_nce_fwd_local",
                 "scp_nce_loss_from_stats", "scp_version",
                 "scp_last_error_string"]:
        assert name in decls, name


def test_library_exports_every_declared_symbol():
    from speechclip_plus_b200 import _lib
    lib = _lib.load()  # builds with nvcc when the .so is missing / stale
    decls = _declared_functions()
    assert len(decls) >= 20
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in scp_b200.h but not exported by libscp_b200.so"
    assert set(_lib.SIGNATURES) == set(decls), set(_lib.SIGNATURES) ^ set(decls)
    for name, nargs in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, binding has {len(_lib.SIGNATURES[name][1])}"


def test_pure_host_entry_points():
    """The handful of functions that do not touch the device."""
    from speechclip_plus_b200 import _lib
    lib = _lib.load()
    assert lib.scp_version() >= 100
    assert lib.scp_last_error_string(0) == b"ok"
    assert b"workspace" in lib.scp_last_error_string(-3)
    assert lib.scp_vq_padded_vocab(49408) == 49408
    assert lib.scp_vq_padded_vocab(8112) == 8192
    assert lib.scp_vq_padded_vocab(19787) == 19968
    assert lib.scp_pack_bytes(3, 128, 512) == 3 * 128 * 512 * 4 + 128 * 8
    assert lib.scp_vq_fwd_workspace_bytes(2048, 49408, 512) > 0
    assert lib.scp_vq_bwd_workspace_bytes(2048, 49408, 512) > 0
    assert lib.scp_nce_workspace_bytes(1024, 512) > 0
    assert lib.scp_wsum_bwd_workspace_bytes(13, 256, 249, 768) > 0
    assert lib.scp_num_launches() >= 0


def test_sass_contains_blackwell_tensor_core_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    from speechclip_plus_b200 import _lib
    _lib.load()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path (tried once as a reduction helper: it stalls behind tcgen05)
