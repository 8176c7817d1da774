"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: batch sharding, max-over-ranks timing, and the
data-parallel gradient all-reduce that is the only collective around the operator."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ceigm_unet_b200 import dist as D
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # 1. shards are disjoint, ordered and cover the batch
    start, stop = D.shard_batch(25, r, w)
    spans = [None] * w
    dist.all_gather_object(spans, (start, stop))
    assert spans[0][0] == 0 and spans[-1][1] == 25
    assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    # 2. timing is the slowest rank's
    assert D.max_over_ranks(1.0 + rank) == float(world)
    # 3. gradient averaging equals the full-batch gradient of a mean loss (what DDP computes)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    x = torch.randn(8, 6)
    full = torch.nn.Linear(6, 3)
    full.load_state_dict(lin.state_dict())
    full(x).square().mean().backward()
    a, b = D.shard_batch(8, r, w)
    lin(x[a:b]).square().mean().backward()
    D.allreduce_module_grads_(lin)
    for p, q in zip(lin.parameters(), full.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6)
    # 4. the overlapped bucketed reducer gives the same averaged gradients, step after step, with an unused parameter
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net.unused = torch.nn.Parameter(torch.ones(7))
    ref_net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ref_net.load_state_dict({k: v for k, v in net.state_dict().items() if k != "unused"})
    red = D.GradReducer(net, bucket_bytes=64)            # tiny buckets: several collectives in flight
    assert len(red.buckets) >= 2
    for step in range(2):
        xb = torch.randn(8, 6, generator=torch.Generator().manual_seed(10 + step))
        ref_net.zero_grad()
        ref_net(xb).square().mean().backward()
        red.zero_grad()
        net(xb[a:b]).square().mean().backward()
        red.finish()
        for (n1, p), (n2, q) in zip(list(net.named_parameters())[1:], ref_net.named_parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), (step, n1)
        assert float(net.unused.grad.abs().sum()) == 0.0
    red.remove()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_batch_edge_cases():
    from ceigm_unet_b200.dist import shard_batch
    assert [shard_batch(24, r, 8) for r in range(8)] == [(3 * r, 3 * r + 3) for r in range(8)]
    assert [shard_batch(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]      # empty shard for the last rank
    assert shard_batch(0, 0, 2) == (0, 0)
