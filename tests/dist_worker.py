"""torchrun worker for tests/test_gpu_dist.py: sharded calc_power on a split catalogue -> npz (rank 0)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))

import cases  # noqa: E402

from abacusutils_b200 import dist as abk_dist  # noqa: E402


def main():
    out, case = sys.argv[1], sys.argv[2]
    reroute = len(sys.argv) > 3 and sys.argv[3] == '1'
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    c = cases.POWER_CASES[case]
    pos, w, pos2, w2 = cases.power_inputs(c)

    def share(a):  # interleaved shares: every rank holds particles from all over the box
        return None if a is None else np.ascontiguousarray(a[rank::world])

    t = abk_dist.calc_power(share(pos), c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'),
                            logk=c['logk'], paste='TSC', nmesh=c['nmesh'], compensated=c['compensated'],
                            interlaced=c['interlaced'], w=share(w), pos2=share(pos2), w2=share(w2), poles=c['poles'],
                            force_reroute=reroute)
    if rank == 0:
        np.savez(out, **{k: np.asarray(t[k]) for k in t.keys()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
