"""Multi-GPU consistency check, launched by tests/test_gpu_dist.py, by __graft_entry__.smoke() when the box
has two GPUs, or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_check.py

The comparison itself lives in fds_b200/selfcheck.py (bench.py --gpus N runs it too): sharded state, both
communication modes, against the same problem on one GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fds_b200.selfcheck import sharded_vs_single  # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    res = sharded_vs_single(local)
    if rank == 0:
        print(json.dumps(res))
        print('DIST_CHECK_OK' if res['pass'] else 'DIST_CHECK_FAILED')
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res['pass'] else 1)


if __name__ == '__main__':
    main()
