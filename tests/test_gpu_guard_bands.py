"""GPU: no kernel writes outside the buffers the C ABI documents.

compute-sanitizer is not available on the GPU pool, so every output and workspace of the entry points below is carved
out of ONE arena with 4 KB sentinel bands in between; after the call all bands must be untouched.  Shapes are chosen to
hit the padding paths (M not a multiple of 128, V not a multiple of 256, odd batch sizes)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

BAND = 4096
FILL = 0xA5


class Arena:
    def __init__(self, nbytes_total: int):
        self.buf = torch.full((nbytes_total,), FILL, dtype=torch.uint8, device="cuda")
        self.off = BAND
        self.regions = []

    def take(self, nbytes: int, dtype=torch.uint8):
        start = (self.off + 255) // 256 * 256
        end = start + nbytes
        assert end + BAND <= self.buf.numel(), "arena too small"
        self.regions.append((start, end))
        self.off = end + BAND
        return self.buf[start:end].view(dtype)

    def assert_bands_intact(self, what: str):
        mask = torch.ones(self.buf.numel(), dtype=torch.bool, device="cuda")
        for s, e in self.regions:
            mask[s:e] = False
        bad = (self.buf[mask] != FILL).nonzero()
        assert bad.numel() == 0, f"{what}: {bad.numel()} bytes written outside the documented buffers"


def _masked(cols):
    return (ctypes.c_int32 * len(cols))(*cols), len(cols)


@pytest.mark.parametrize("shape", [(200, 8, 1000, 64), (2048 - 8, 8, 8112, 512), (24, 3, 49408, 512), (130, 5, 3000, 768)])
def test_vq_fwd_bwd_stay_inside_their_buffers(shape):
    from speechclip_plus_b200 import _lib
    lib = _lib.load()
    M, K, V, D = shape
    gen = torch.Generator(device="cuda").manual_seed(M + V)
    table = torch.randn(V, D, device="cuda", generator=gen) * 0.02
    kw = torch.randn(M, D, device="cuda", generator=gen) * 0.02
    g = torch.randn(M, D, device="cuda", generator=gen)
    tau = torch.tensor([0.1], device="cuda")
    Vp = int(lib.scp_vq_padded_vocab(V))
    Mp = (M + 127) // 128 * 128
    stream = _lib.stream_ptr(torch.device("cuda"))
    ws_f, ws_b = lib.scp_vq_fwd_workspace_bytes(M, V, D), lib.scp_vq_bwd_workspace_bytes(M, V, D)
    ar = Arena(ws_f + ws_b + 3 * Vp * D * 2 + 3 * Mp * D * 4 + 64 * Vp + 2 ** 20)
    hat = ar.take(Vp * D * 2, torch.float16)
    hat_t = ar.take(D * Vp * 2, torch.float16)
    norm = ar.take(Vp * 4, torch.float32)
    mean = ar.take((D + 1) * 4, torch.float32)
    _lib.check(lib.scp_vq_prepare_table(_lib.ptr(table), V, D, _lib.ptr(hat), _lib.ptr(hat_t), _lib.ptr(norm),
                                        _lib.ptr(mean), stream), "prepare_table")
    idx = ar.take(M * 8, torch.int64)
    kw_out = ar.take(M * D * 4, torch.float32)
    row_stats = ar.take(M * 4 * 4, torch.float32)
    code_hist = ar.take(Vp * 4, torch.float32)
    avg_probs = ar.take(Vp * 4, torch.float32)
    metrics = ar.take((3 + K) * 4, torch.float32)
    kw_hat = ar.take(Mp * D * 2, torch.float16)
    wsf = ar.take(ws_f)
    masked, n_masked = _masked([0, 2, 3])
    _lib.check(lib.scp_vq_fwd(_lib.ptr(kw), M, K, V, D, _lib.ptr(hat), _lib.ptr(norm), _lib.ptr(table), masked, n_masked,
                              _lib.ptr(tau), _lib.ptr(idx), _lib.ptr(kw_out), _lib.ptr(row_stats), _lib.ptr(code_hist),
                              _lib.ptr(avg_probs), _lib.ptr(metrics), _lib.ptr(kw_hat), _lib.ptr(wsf), ws_f, stream),
               "scp_vq_fwd")
    g_kw = ar.take(M * D * 4, torch.float32)
    g_tau = ar.take(4, torch.float32)
    wsb = ar.take(ws_b)
    _lib.check(lib.scp_vq_bwd(_lib.ptr(g), _lib.ptr(kw), M, V, D, _lib.ptr(kw_hat), _lib.ptr(hat), _lib.ptr(hat_t),
                              _lib.ptr(norm), _lib.ptr(mean), _lib.ptr(row_stats), masked, n_masked, _lib.ptr(tau),
                              _lib.ptr(g_kw), _lib.ptr(g_tau), _lib.ptr(wsb), ws_b, stream), "scp_vq_bwd")
    torch.cuda.synchronize()
    ar.assert_bands_intact(f"vq M={M} V={V} D={D}")
    assert int(idx.min()) >= 0 and int(idx.max()) < V and torch.isfinite(g_kw).all()
    assert abs(float(avg_probs[:V].sum()) - 1.0) < 1e-3 and float(code_hist.sum()) == M


@pytest.mark.parametrize("shape", [(13, 3, 50, 768), (25, 2, 33, 1024), (5, 7, 9, 64)])
@pytest.mark.parametrize("norm_mode", [0, 1, 2, 3])
def test_wsum_stays_inside_its_buffers(shape, norm_mode):
    from speechclip_plus_b200 import _lib
    lib = _lib.load()
    L, B, T, D = shape
    gen = torch.Generator(device="cuda").manual_seed(L * B)
    layers = [torch.randn(T, B, D, device="cuda", generator=gen).transpose(0, 1) for _ in range(L)]
    w = torch.randn(L, device="cuda", generator=gen)
    stream = _lib.stream_ptr(torch.device("cuda"))
    ws_b = lib.scp_wsum_bwd_workspace_bytes(L, B, T, D)
    ar = Arena(3 * B * T * D * 4 + L * B * T * D * 4 + ws_b + 2 ** 20)
    y = ar.take(B * T * D * 4, torch.float32)
    us = ar.take(L * B * 4, torch.float32)
    ptrs = _lib.ptr_array(layers)
    v0 = layers[0]
    if norm_mode == 3:
        _lib.check(lib.scp_wsum_utt_scale(ptrs, L, B, T, D, v0.stride(0), v0.stride(1), 0, _lib.ptr(us), stream), "utt_scale")
    _lib.check(lib.scp_wsum_fwd(ptrs, L, B, T, D, v0.stride(0), v0.stride(1), 0, _lib.ptr(w), norm_mode, 1e-5,
                                _lib.ptr(us), _lib.ptr(y), 0, stream), "wsum_fwd")
    gy = torch.randn(B, T, D, device="cuda", generator=gen)
    dw = ar.take(L * 4, torch.float32)
    ws = ar.take(ws_b)
    g_layers = None
    if norm_mode != 3:
        gl = [ar.take(B * T * D * 4, torch.float32) for _ in range(L)]
        g_layers = _lib.ptr_array(gl)
    _lib.check(lib.scp_wsum_bwd(ptrs, L, B, T, D, v0.stride(0), v0.stride(1), 0, _lib.ptr(w), norm_mode, 1e-5,
                                _lib.ptr(us), _lib.ptr(gy), 0, _lib.ptr(dw),
                                g_layers if g_layers is not None else ctypes.cast(None, ctypes.POINTER(ctypes.c_void_p)),
                                _lib.ptr(ws), ws_b, stream), "wsum_bwd")
    torch.cuda.synchronize()
    ar.assert_bands_intact(f"wsum {shape} mode {norm_mode}")
    assert torch.isfinite(y).all() and torch.isfinite(dw).all()


@pytest.mark.parametrize("cfg", [(37, 64, 256, 0, 37), (300, 128, 40, 60, 100), (1000, 512, 96, 904, 1000)])
def test_nce_stays_inside_its_buffers(cfg):
    from speechclip_plus_b200 import _lib
    lib = _lib.load()
    N, D, _, lo, hi = cfg
    gen = torch.Generator(device="cuda").manual_seed(N)
    a = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=gen), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=gen), dim=-1)
    ids = torch.randint(0, max(N // 4, 1), (N,), device="cuda", generator=gen)
    ls = torch.tensor([2.659], device="cuda")
    stream = _lib.stream_ptr(torch.device("cuda"))
    wsn = lib.scp_nce_workspace_bytes(N, D)
    ar = Arena(wsn + 4 * N * D * 4 + 2 ** 20)
    loss = ar.take(4, torch.float32)
    lse_r = ar.take(N * 4, torch.float32)
    lse_c = ar.take(N * 4, torch.float32)
    ws = ar.take(wsn)
    _lib.check(lib.scp_nce_fwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ids), N, D, _lib.ptr(ls), 0.0, 0.0, 0, 1, 1, 1,
                               _lib.ptr(loss), _lib.ptr(lse_r), _lib.ptr(lse_c), _lib.ptr(ws), wsn, stream), "nce_fwd")
    g = torch.ones(1, device="cuda")
    dA = ar.take((hi - lo) * D * 4, torch.float32)
    dB = ar.take((hi - lo) * D * 4, torch.float32)
    dT = ar.take(4, torch.float32)
    _lib.check(lib.scp_nce_bwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ids), N, D, _lib.ptr(ls), 0.0, 0.0, 0, 1, 1,
                               _lib.ptr(lse_r), _lib.ptr(lse_c), _lib.ptr(g), lo, hi, 1, _lib.ptr(dA), _lib.ptr(dB),
                               _lib.ptr(dT), _lib.ptr(ws), wsn, stream), "nce_bwd")
    torch.cuda.synchronize()
    ar.assert_bands_intact(f"nce N={N} D={D} rows [{lo},{hi})")
    assert torch.isfinite(loss).all() and torch.isfinite(dA).all() and torch.isfinite(dB).all()


def test_cif_and_splice_stay_inside_their_buffers():
    from speechclip_plus_b200 import _lib
    lib = _lib.load()
    stream = _lib.stream_ptr(torch.device("cuda"))
    gen = torch.Generator(device="cuda").manual_seed(3)
    B, S, C = 7, 61, 192
    x = torch.randn(B, S, C, device="cuda", generator=gen)
    alpha = torch.rand(B, S, device="cuda", generator=gen) * 0.4
    ar = Arena(8 * B * S * C * 4 + 2 ** 20)
    csum = ar.take(B * S * 4, torch.float32)
    flen = ar.take(B * 8, torch.int64)
    _lib.check(lib.scp_cif_plan(_lib.ptr(alpha), B, S, 1.0, 75, _lib.ptr(csum), _lib.ptr(flen), stream), "cif_plan")
    T = int(flen.max())
    out = ar.take(B * (T + 1) * C * 4, torch.float32)
    fm = ar.take(B * S, torch.uint8)
    tw = ar.take(B * 4, torch.float32)
    _lib.check(lib.scp_cif_fire_fwd(_lib.ptr(x), _lib.ptr(alpha), _lib.ptr(csum), _lib.ptr(flen), B, S, C, 1.0, T,
                                    _lib.ptr(out), _lib.ptr(fm), _lib.ptr(tw), stream), "cif_fire_fwd")
    fnew = ar.take(B * 8, torch.int64)
    _lib.check(lib.scp_cif_tail(_lib.ptr(out), B, T + 1, C, _lib.ptr(flen), _lib.ptr(tw), 1.0, 0.5, 75, _lib.ptr(fnew),
                                stream), "cif_tail")
    T2 = int(fnew.max())
    g = torch.randn(B, T2, C, device="cuda", generator=gen)
    gx = ar.take(B * S * C * 4, torch.float32)
    ga = ar.take(B * S * 4, torch.float32)
    ws = ar.take(2 * B * S * 4)
    _lib.check(lib.scp_cif_fire_bwd(_lib.ptr(g), T2, _lib.ptr(x), _lib.ptr(alpha), _lib.ptr(csum), B, S, C, 1.0, T,
                                    _lib.ptr(fnew), _lib.ptr(flen), _lib.ptr(tw), 0.5, _lib.ptr(gx), _lib.ptr(ga),
                                    _lib.ptr(ws), 2 * B * S * 4, stream), "cif_fire_bwd")
    # splice
    Bk, Kmax, D, V, L = 5, 9, 64, 300, 77
    table = torch.randn(V, D, device="cuda", generator=gen)
    pos = torch.randn(L, D, device="cuda", generator=gen)
    kw = torch.randn(Bk, Kmax, D, device="cuda", generator=gen)
    num = torch.tensor([9, 0, 3, 20, 7], device="cuda")  # 20 > Kmax: clamped
    xs = ar.take(Bk * L * D * 4, torch.float32)
    eot = ar.take(Bk * 8, torch.int64)
    _lib.check(lib.scp_kw_splice_fwd(_lib.ptr(kw), _lib.ptr(num), 0, _lib.ptr(table), _lib.ptr(pos), 0, Bk, Kmax, D, L,
                                     V - 2, V - 1, _lib.ptr(xs), _lib.ptr(eot), stream), "splice_fwd")
    gk = ar.take(Bk * Kmax * D * 4, torch.float32)
    _lib.check(lib.scp_kw_splice_bwd(_lib.ptr(xs), 0, _lib.ptr(num), 0, Bk, Kmax, D, L, _lib.ptr(gk), stream), "splice_bwd")
    torch.cuda.synchronize()
    ar.assert_bands_intact("cif + splice")
    assert eot.tolist() == [10, 1, 4, 10, 8]
