"""world_size-2 gloo tests (CPU) of the host logic behind the N > 1 path:
  * the SI prototype exchange: all-reducing per-class SUMS and COUNTS before the division reproduces, on every rank,
    the bank a single process computes on the global batch (oracle arithmetic);
  * bucketed gradient averaging over an arena equals one whole-arena average, with the engine's stage ordering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vqacl_b200.modeling import plan_grad_buckets
        g = torch.Generator().manual_seed(7)
        B, d = 12, 768
        hidden = torch.randn(B, 26, d, generator=g)
        lab = torch.zeros(B, 10)
        lab[torch.arange(B), torch.randint(0, 4, (B,), generator=g)] = 1
        full, cnt_full = O.calculate_current_prototype(hidden[:, :20], lab)
        # this rank's shard -> un-divided sums + counts (what vqacl_proto_sums writes), all-reduce, then divide
        sl = slice(rank * B // world, (rank + 1) * B // world)
        m = hidden[sl, :20].mean(1)
        sums = lab[sl].t() @ m
        cnt = lab[sl].sum(0)
        dist.all_reduce(sums)
        dist.all_reduce(cnt)
        proto = sums / torch.where(cnt <= 0, torch.ones_like(cnt), cnt)[:, None]
        ok1 = torch.equal(cnt, cnt_full) and torch.allclose(proto, full, rtol=1e-5, atol=1e-6)

        # bucketed averaging == whole-arena averaging
        ranges = [(0, 0), (300, 400), (200, 300), (100, 200), (400, 520), (800, 900), (700, 800), (520, 700), (900, 1000)]
        arena = torch.randn(1000, generator=torch.Generator().manual_seed(100 + rank))
        whole = arena.clone()
        dist.all_reduce(whole)
        whole /= world
        plan = plan_grad_buckets(ranges, 150)
        for s in range(len(ranges)):
            for a, b in plan.get(s, ()):
                dist.all_reduce(arena[a:b])
                arena[a:b] /= world
        ok2 = torch.allclose(arena[100:], whole[100:], rtol=0, atol=1e-7)
        # sharded step tail == replicated step tail: reduce-scatter (emulated: gloo has none) + AdamW on the owned slices +
        # all-gather, tail all-reduced and updated everywhere, global norm from owned slices (+ tail once)
        from vqacl_b200.modeling import owned_slices, plan_shards
        n_tail, n_train = 960, 1040
        sranges = [(0, 0), (320, 480), (160, 320), (0, 160), (480, 640), (800, 960), (640, 800), (960, 1040)]
        _, buckets, tail = plan_shards(sranges, n_tail, n_train, 200, world)
        gen = torch.Generator().manual_seed(200 + rank)
        g_local = torch.randn(n_train, generator=gen)
        p0 = torch.randn(n_train, generator=torch.Generator().manual_seed(5))
        g_avg = g_local.clone()
        dist.all_reduce(g_avg)
        g_avg /= world

        def adam(p, g, coef):          # first HF-AdamW step (m = v = 0), lr 0.1
            g = g * coef
            m, v = 0.1 * g, 0.001 * g * g
            return p - 0.1 * (0.001 ** 0.5 / 0.1) * m / (v.sqrt() + 1e-6)
        coef_full = min(1.0, 5.0 / (g_avg.norm().item() + 1e-6))
        want = adam(p0, g_avg, coef_full)
        own = owned_slices(buckets, world, rank)
        ss = sum(float((g_avg[a:b] ** 2).sum()) for a, b in own) + (float((g_avg[tail[0]:tail[1]] ** 2).sum()) if rank == 0 else 0.0)
        t = torch.tensor([ss], dtype=torch.float64)
        dist.all_reduce(t)
        coef = min(1.0, 5.0 / (t.item() ** 0.5 + 1e-6))
        p = p0.clone()
        for a, b in own:
            p[a:b] = adam(p0[a:b], g_avg[a:b], coef)
        p[tail[0]:tail[1]] = adam(p0[tail[0]:tail[1]], g_avg[tail[0]:tail[1]], coef)
        for (a, b), (oa, ob) in zip(buckets, own):
            parts = [torch.empty(ob - oa) for _ in range(world)]
            dist.all_gather(parts, p[oa:ob].clone())
            p[a:b] = torch.cat(parts)
        ok3 = abs(coef - coef_full) < 1e-6 and torch.allclose(p, want, rtol=1e-5, atol=1e-6)
        q.put((rank, ok1, ok2 and ok3))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "prototype sums/counts exchange does not reproduce the global-batch bank"
    assert all(r[2] for r in res), "bucketed gradient averaging differs from whole-arena averaging"
