"""Two-GPU data-parallel parity (needs >= 2 CUDA devices; skipped otherwise): two ranks with half the batch each, NCCL
gradient averaging from inside backward and the SI prototype sum/count exchange, against ONE process on the global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from helpers import O, cos, make_pair, rel_err
        import vqacl_b200 as V
        B = 16
        full = O.synthetic_batch(B, seed=321, task_id=2, rehearsal=True)
        shard = {k: v[rank * B // world:(rank + 1) * B // world].clone() for k, v in full.items()}
        _, m = make_pair(layers=2, device=f"cuda:{rank}")
        m.train()
        m.grad_bucket_elems = 1 << 20          # several buckets even for the 2-layer model
        g = torch.Generator().manual_seed(5)
        m.Q_prototype = torch.randn(10, 768, generator=g)
        # the reference wraps the model in DDP and then calls model.module.train_step (vqacl.py:127-129,438): the wrapper's
        # reducer never fires; gradient averaging happens inside loss.backward()
        ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[rank], find_unused_parameters=True)
        r = ddp.module.train_step(shard, 2, 0.5, 0.3)
        r["loss"].backward()
        torch.cuda.synchronize()
        ok = True
        msg = ""
        if rank == 0:
            _, ref = make_pair(layers=2, device="cuda:0")
            ref.train()
            ref.sync_grads = ref.sync_prototypes = False
            g = torch.Generator().manual_seed(5)
            ref.Q_prototype = torch.randn(10, 768, generator=g)
            rr = ref.train_step(full, 2, 0.5, 0.3)
            rr["loss"].backward()
            torch.cuda.synchronize()
            n = m._engine.n_train
            c = cos(m._engine.G[:n], ref._engine.G[:n])
            e = rel_err(m._engine.G[:n], ref._engine.G[:n])
            pq = rel_err(m.Q_prototype, ref.Q_prototype)
            pv = rel_err(m.V_prototype, ref.V_prototype)
            cnt = torch.equal(m.V_prototype_num, ref.V_prototype_num) and torch.equal(m.Q_prototype_num, ref.Q_prototype_num)
            ok = c > 0.999 and e < 3e-2 and pq < 1e-5 and pv < 1e-5 and cnt
            msg = f"grad cos {c:.6f} rel {e:.3e} Qbank {pq:.2e} Vbank {pv:.2e} counts {cnt}"
        # both ranks must hold identical (averaged) gradients and banks
        n = m._engine.n_train
        mine = torch.stack([m._engine.G[:n].double().sum(), m.V_prototype.double().sum(), m.Q_prototype.double().sum()])
        other = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        same = all(torch.equal(o, other[0]) for o in other)
        q.put((rank, ok and same, msg + f" identical-across-ranks {same}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_equals_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"


def _worker_sharded(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from helpers import O, cos, make_pair
        import vqacl_b200 as V
        B, steps = 16, 3
        _, m = make_pair(layers=2, device=f"cuda:{rank}")
        m.train()
        m.grad_bucket_elems = 1 << 20
        before = {k: v.clone() for k, v in m.state_dict().items()}
        opt = V.FusedAdamW(m, lr=1e-3, overlap_with_next_forward=True)        # shard_state=None -> sharded (2 ranks)
        assert opt.shard and m.shard_optimizer and opt.exp_avg.numel() < 0.75 * m._engine.n_train
        for i in range(steps):
            full = O.synthetic_batch(B, seed=500 + i, task_id=1, rehearsal=(i == 2))
            shard = {k: v[rank * B // world:(rank + 1) * B // world].clone() for k, v in full.items()}
            m.train_step(shard, 1, 0.5, 0.3)["loss"].backward()
            opt.step(max_grad_norm=5.0)
            opt.zero_grad()
        toks = m.test_step(O.synthetic_batch(8, seed=9))["token_ids"]       # generate right after an overlapped sharded step
        after = m.state_dict()                                              # all-gathers the fp32 masters
        ok, msg = True, ""
        if rank == 0:
            _, ref = make_pair(layers=2, device="cuda:0")
            ref.train()
            ref.sync_grads = ref.sync_prototypes = False
            ropt = V.FusedAdamW(ref, lr=1e-3, shard_state=False)
            for i in range(steps):
                full = O.synthetic_batch(B, seed=500 + i, task_id=1, rehearsal=(i == 2))
                ref.train_step(full, 1, 0.5, 0.3)["loss"].backward()
                ropt.step(max_grad_norm=5.0)
                ropt.zero_grad()
            rafter = ref.state_dict()
            worst = (1.0, "")
            for k in after:
                du, dr = (after[k] - before[k]).float(), (rafter[k] - before[k]).float()
                if dr.abs().max() == 0:
                    continue
                c = cos(du, dr)
                if c < worst[0]:
                    worst = (c, k)
            rtoks = ref.test_step(O.synthetic_batch(8, seed=9))["token_ids"]
            gn = abs(opt.grad_sumsq.item() - ropt.grad_sumsq.item()) / ropt.grad_sumsq.item()
            ok = worst[0] > 0.99 and gn < 1e-3 and torch.equal(toks.cpu(), rtoks.cpu())
            msg = f"worst update cosine {worst[0]:.5f} ({worst[1]}), grad-norm^2 rel diff {gn:.2e}, same answers {torch.equal(toks.cpu(), rtoks.cpu())}"
        mine = torch.stack([v.double().sum() for v in after.values()])
        other = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        same = all(torch.equal(o, other[0]) for o in other)
        q.put((rank, ok and same, msg + f" identical-weights-across-ranks {same}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_optimizer_equals_single_process():
    """Reduce-scattered gradients + AdamW on each rank's slices + all-gathered bf16 weights (overlapped with the next forward)
    leave every rank with the weights ONE process reaches on the global batch with the replicated optimizer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"
    print("2-GPU sharded step:", [m for _, _, m in res])
