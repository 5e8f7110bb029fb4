"""Ring-of-ranks plumbing for the N-sharded shadowing loop (one process per GPU).

The state dimension is cut into contiguous blocks, one per rank, exactly as the
reference shards it over MPI ranks (fds/compute.py:87-89) and as its Lorenz-96 mock
solver does (tests/solvers/mock_fun3d/fun3d:15-37).  Three exchanges exist:

* halo exchange for the i-2..i+1 stencil: a rank sends its last ``8*h`` sites to the
  right neighbour and its first ``4*h`` sites to the left one (valid for ``h`` fused
  RK4 steps) -- ``torch.distributed`` send/recv (NCCL over NVLink on GPUs);
* one all-reduce (sum) of the (m+2)^2 extended Gram per CholeskyQR pass;
* an all-gather of the per-rank objective partial sums followed by a FIXED-ORDER
  binary-tree sum over ranks, so J does not depend on the number of ranks when the
  shard size is a power of two.

Everything here works on any ``torch`` tensors, so the same code is exercised on
CPU tensors with the ``gloo`` backend in the tests (world_size 2) and on aliased
device buffers with ``nccl`` on the GPU box.
"""
import numpy as np


def shard_range(n_global, rank, world):
    """Contiguous block of sites owned by ``rank`` (same split as mpi_range,
    reference fds/compute.py:87-89, for world dividing n_global)."""
    base, rem = divmod(int(n_global), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tree_sum_leading(parts):
    """Balanced binary tree over the leading axis, zero-padded to a power of two,
    lower index on the left -- the rank-level continuation of the per-shard tree of
    fds_b200/csrc/l96_core.cuh (and oracle tree_sum)."""
    parts = [np.asarray(p, dtype=np.float64) for p in parts]
    n = 1
    while n < len(parts):
        n *= 2
    parts = parts + [np.zeros_like(parts[0])] * (n - len(parts))
    while len(parts) > 1:
        parts = [parts[i] + parts[i + 1] for i in range(0, len(parts), 2)]
    return parts[0]


class Ring:
    """Periodic ring of ``world`` ranks over a torch.distributed process group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world

    def _global(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def exchange_halo(self, send_right, send_left, recv_left, recv_right):
        """send_right -> right neighbour's recv_left ; send_left -> left neighbour's recv_right."""
        d = self.dist
        ops = [d.P2POp(d.isend, send_right, self._global(self.right), self.group),
               d.P2POp(d.isend, send_left, self._global(self.left), self.group),
               d.P2POp(d.irecv, recv_left, self._global(self.left), self.group),
               d.P2POp(d.irecv, recv_right, self._global(self.right), self.group)]
        for w in d.batch_isend_irecv(ops):
            w.wait()

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allgather(self, t):
        import torch
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return out

    def ordered_sum(self, t):
        """Fixed-order tree sum over ranks of a (small) tensor; returns a host array,
        identical on all ranks."""
        return tree_sum_leading([p.cpu().numpy() for p in self.allgather(t)])
