"""torchrun worker of tests/test_gpu_configs.py::test_config5_gathered_two_gpus.

Every rank extracts the features of its shard of seeded clips on its GPU, the blocks are gathered on rank 0 over
NCCL and, second, written by the kernels straight into rank 0's peer-mapped buffer (pyaudioanalysis_b200.dist), and rank 0 checks the gathered tensor against the oracle clip by clip -- including an
uneven split (7 clips over the ranks) so the padding / trimming of the collective is covered.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import st_oracle as O                              # noqa: E402  (checker only)
from tests.parity import check_features                        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pyaudioanalysis_b200.dist import feature_extraction_sharded
    n_clips, n, fs, w, s = 7, 48000, 16000, 800, 400

    def clips(lo, hi):
        return torch.from_numpy(np.stack([O.synth_clip(500 + i, n, fs) for i in range(lo, hi)])).cuda()

    refs = [O.feature_extraction(O.synth_clip(500 + i, n, fs), fs, w, s)[0] for i in range(n_clips)] if rank == 0 else None
    for mode in ("nccl", "p2p", "p2p_store"):      # collective baseline; copy-engine push / kernel stores into the root's peer-mapped buffer
        got = feature_extraction_sharded(clips, n_clips, fs, w, s, deltas=True, gather_to=0, gather=mode)
        if rank == 0:
            assert got.shape == (n_clips, 68, (n - w) // s + 1), got.shape
            got = got.cpu().numpy()
            for i in range(n_clips):
                check_features(got[i], refs[i], w // 2, "%s-gathered clip %d" % (mode, i))
        else:
            assert got is None
        dist.barrier()
    if rank == 0:
        print("DIST_GPU_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
