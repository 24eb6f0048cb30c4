"""CPU, world_size 2, gloo: the N>1 plumbing of the loss (pack layout -> all-gather -> unpack -> global loss with
local-row gradients).  The CUDA kernels cannot run here; the per-rank packed buffers and the loss are produced by the
oracle, the collective and (un)packing code under test is the product's (speechclip_plus_b200/model/kw_glue.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, result_dir: str):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import speechclip_oracle as oracle
        from speechclip_plus_b200.model import kw_glue
        n, D = 6, 16
        N = n * world
        gen = torch.Generator().manual_seed(1234)  # every rank builds the same global problem
        audio = torch.randn(N, D, generator=gen, dtype=torch.float64)
        image = torch.randn(N, D, generator=gen, dtype=torch.float64)
        ids = torch.randint(0, 5, (N,), generator=gen)
        r0, r1 = kw_glue.shard_rows(n, rank)
        assert (r0, r1) == (rank * n, (rank + 1) * n)
        # per-rank pack in the layout of scp_l2norm_pack: [feat blocks (n,D) f32 ..., ids (n,) i64]
        a_loc = audio[r0:r1].clone().requires_grad_(True)
        a_hat = oracle.l2_normalise(a_loc)
        i_hat = oracle.l2_normalise(image[r0:r1])
        packed = torch.cat([i_hat.float().contiguous().view(torch.uint8).reshape(-1),
                            a_hat.detach().float().contiguous().view(torch.uint8).reshape(-1),
                            ids[r0:r1].contiguous().view(torch.uint8).reshape(-1)])
        assert packed.numel() == kw_glue.pack_nbytes(2, n, D)
        gathered = kw_glue.all_gather_packed(packed)
        assert gathered.shape == (world, packed.numel())
        (g_img, g_aud), g_ids = kw_glue.unpack_gathered(gathered, 2, n, D)
        assert torch.equal(g_ids, ids)
        assert torch.allclose(g_img.double(), oracle.l2_normalise(image), atol=1e-6)
        assert torch.allclose(g_aud.double(), oracle.l2_normalise(audio), atol=1e-6)
        # global loss evaluated redundantly on every rank; gradient only into the local rows
        aud_full = g_aud.double().clone()
        aud_full[r0:r1] = a_hat  # local rows keep their autograd history, remote rows are constants
        loss = oracle.nce_forward(aud_full, g_img.double(), g_ids, 1 / 0.07)
        (g_local,) = torch.autograd.grad(loss, [a_loc])
        # single-process reference: gradient of the same global loss w.r.t. all rows
        a_all = audio.clone().requires_grad_(True)
        ref = oracle.nce_forward(oracle.l2_normalise(a_all), oracle.l2_normalise(image), ids, 1 / 0.07)
        (g_all,) = torch.autograd.grad(ref, [a_all])
        assert abs(loss.item() - ref.item()) < 1e-6
        assert torch.allclose(g_local, g_all[r0:r1], atol=1e-6)
        # DDP semantics: a parameter shared by all rows (here: a global scale s, feature = s * x) gets
        # sum_r (local-row gradient); DDP averages, so the loss is scaled by world_size (kw_glue.ddp_grad_scale)
        p_grad_local = (g_local * audio[r0:r1]).sum().reshape(1)
        summed = p_grad_local.clone()
        dist.all_reduce(summed)
        averaged_scaled = summed / world * kw_glue.ddp_grad_scale(world)
        p_ref = (g_all * audio).sum()
        assert abs(averaged_scaled.item() - p_ref.item()) < 1e-6 * max(1.0, abs(p_ref.item()))
        # ---- sharded forward (SURVEY section 8(e) option B): each rank evaluates the denominators of its own rows and
        #      columns (losses.py:224-243 per shard), the ranks exchange (3, n) statistics with the product's gather
        from speechclip_plus_b200.module import losses as scp_losses
        A, Bm = oracle.l2_normalise(audio), oracle.l2_normalise(image)
        scale = 1 / 0.07
        S = A @ Bm.t() * scale
        mask = oracle.nce_mask(ids, N)
        neg_inf = torch.full_like(S, float("-inf"))
        lse_row_loc = torch.logsumexp(torch.where(mask, S, neg_inf)[r0:r1], dim=1)        # local rows of A vs all B
        lse_col_loc = torch.logsumexp(torch.where(mask, S, neg_inf)[:, r0:r1], dim=0)     # local rows of B vs all A
        pos_loc = (A[r0:r1] * Bm[r0:r1]).sum(-1)
        stats = torch.stack([lse_row_loc, lse_col_loc, pos_loc]).float()
        stats_all = scp_losses.all_gather_stats(stats, world)
        assert stats_all.shape == (world, 3, n)
        lr = stats_all[:, 0].reshape(-1).double()
        lc = stats_all[:, 1].reshape(-1).double()
        ps = stats_all[:, 2].reshape(-1).double() * scale
        loss_sharded = 0.5 * ((lr - ps).mean() + (lc - ps).mean())
        assert abs(loss_sharded.item() - ref.item()) < 1e-5
        os.environ["SCP_NCE_SHARD_MIN_N"] = "0"
        assert scp_losses._use_sharded_forward(N, r0, r1, None)
        assert not scp_losses._use_sharded_forward(N, 0, N, None)
        assert not scp_losses._use_sharded_forward(N, r0, r1, kw_glue.LOCAL)
        os.environ["SCP_NCE_SHARD_MIN_N"] = "2048"
        assert not scp_losses._use_sharded_forward(N, r0, r1, None)       # small global batches stay redundant
        # ---- CIF quantity loss on the multi-rank path: value = global mean, gradient = reference's after the
        #      ddp_grad_scale * mean-over-ranks recipe (kwClip.py:1031-1038)
        w = torch.tensor(0.7, dtype=torch.float64, requires_grad=True)     # a parameter shared by every rank
        xq = torch.randn(N, generator=gen, dtype=torch.float64)
        tgt = torch.randn(N, generator=gen, dtype=torch.float64)
        q_local = torch.nn.functional.l1_loss(w * xq[r0:r1], tgt[r0:r1])
        q = kw_glue.global_mean_with_local_grad(q_local)
        q_ref = torch.nn.functional.l1_loss(w * xq, tgt)
        assert abs(q.item() - q_ref.item()) < 1e-12
        (gw,) = torch.autograd.grad(q * kw_glue.ddp_grad_scale(world), [w])
        dist.all_reduce(gw)
        gw = gw / world                                                     # DDP averages
        (gw_ref,) = torch.autograd.grad(q_ref, [w])
        assert abs(gw.item() - gw_ref.item()) < 1e-12
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world2_gloo_gather_and_local_row_gradients(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
