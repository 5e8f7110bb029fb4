"""CPU (gloo, world_size 2 and 3): the ring plumbing of the N-sharded shadowing loop --
halo exchange, all-reduce, fixed-order objective sum -- on CPU tensors, driving a numpy
emulation of the sharded Lorenz-96 step (the oracle's step applied to ghost-padded shards).
The same fds_b200.ring code runs over NCCL on device buffers on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fds_b200.ring import Ring, shard_range, tree_sum_leading
from oracle import plugins


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rhs_padded(x, F):
    """l96_rhs_factored on a ghost-padded 1-D shard: valid for indices 2 .. len-2."""
    out = np.full_like(x, np.nan)
    out[2:-1] = (x[3:] - x[:-3]) * x[1:-2] - x[2:-1] + F
    return out


def _euler_sharded(ring, u_local, F, steps, hs):
    """Forward Euler with a halo refreshed every `hs` steps: 2*hs ghosts left, hs right
    (one RHS evaluation per step; the RK4 kernel uses 8 / 4 per step the same way)."""
    n = len(u_local)
    gl, gr = 2 * hs, hs
    buf = torch.zeros(gl + n + gr, dtype=torch.float64)
    buf[gl:gl + n] = torch.from_numpy(u_local)
    t = 0
    while t < steps:
        k = min(hs, steps - t)
        ring.exchange_halo(buf[n:gl + n], buf[gl:gl + gr], buf[0:gl], buf[gl + n:])
        x = buf.numpy().copy()
        for j in range(k):
            dx = _rhs_padded(x, F)
            x = x + 0.01 * dx          # cells whose stencil left the valid halo turn NaN, never read
        buf[gl:gl + n] = torch.from_numpy(x[gl:gl + n])
        t += k
    return buf[gl:gl + n].numpy().copy()


def _worker(rank, world, port, N, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ring = Ring()
        assert (ring.left, ring.right) == ((rank - 1) % world, (rank + 1) % world)
        lo, hi = shard_range(N, rank, world)
        u0 = np.random.RandomState(0).rand(N) * 4 - 1
        # ---- halo exchange drives a sharded solver to the same bits as the global one ----
        got = _euler_sharded(ring, u0[lo:hi], 8.1, 7, 3)
        x = u0.copy()
        for _ in range(7):
            x = x + 0.01 * ((np.roll(x, -1) - np.roll(x, 2)) * np.roll(x, 1) - x + 8.1)
        ok_state = np.array_equal(got, x[lo:hi])
        # ---- Gram all-reduce ----
        g = torch.from_numpy(np.outer(u0[lo:hi][:5], u0[lo:hi][:5]).copy())
        ring.allreduce_sum(g)
        ref = sum(np.outer(u0[a:b][:5], u0[a:b][:5]) for a, b in (shard_range(N, r, world) for r in range(world)))
        ok_gram = np.allclose(g.numpy(), ref, rtol=1e-14)
        # ---- fixed-order objective sum: per-shard canonical tree + rank tree == global tree ----
        part = torch.from_numpy(np.array([plugins.tree_sum(x[lo:hi]), plugins.tree_sum(x[lo:hi] ** 2)]))
        tot = ring.ordered_sum(part)
        ok_sum = True
        if world & (world - 1) == 0 and (N // world) & (N // world - 1) == 0:
            ok_sum = tot[0] == plugins.tree_sum(x) and tot[1] == plugins.tree_sum(x * x)
        # ---- bootstrap of the library's own NCCL communicator (fds_comm_unique_id on rank 0, handed to every
        # rank through the process group -- the host side of Engine(comm='nccl'); the id is opaque bytes) ----
        from fds_b200.device import comm_unique_id
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=ring._global(0), group=ring.group)
        chk = torch.tensor([float(sum(ids[0]))], dtype=torch.float64)
        ring.allreduce_sum(chk)
        ok_id = len(ids[0]) == 128 and chk.item() == world * float(sum(ids[0]))
        q.put((rank, ok_state, ok_gram, ok_sum and ok_id, tot.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,N', [(2, 64), (2, 50), (3, 48)])
def test_ring_two_and_three_ranks(world, N):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tots = [r[4] for r in res]
    assert all(t == tots[0] for t in tots)           # identical on every rank
    for rank, ok_state, ok_gram, ok_sum, _ in res:
        assert ok_state, 'sharded state differs from the global run (rank %d)' % rank
        assert ok_gram and ok_sum


def test_shard_range_covers_everything():
    for N, W in ((40, 4), (37, 3), (1 << 20, 8), (5, 8)):
        edges = [shard_range(N, r, W) for r in range(W)]
        assert edges[0][0] == 0 and edges[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


def test_tree_sum_leading_is_the_rank_level_continuation():
    rng = np.random.RandomState(3)
    x = rng.randn(4096)
    for W in (2, 4, 8):
        parts = [np.array(plugins.tree_sum(x[i * 4096 // W:(i + 1) * 4096 // W])) for i in range(W)]
        assert tree_sum_leading(parts) == plugins.tree_sum(x)
    parts = [np.array(float(i)) for i in range(3)]
    assert tree_sum_leading(parts) == (0.0 + 1.0) + (2.0 + 0.0)
