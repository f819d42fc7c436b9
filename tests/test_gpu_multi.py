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
