"""worker for tests/test_dist_gloo.py: python tests/_dist_worker.py RANK WORLD PORT NUM_FRAMES OUT.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from detzero_b200 import dist as dz  # noqa: E402

rank, world, port, num_frames, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5]
os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group('gloo', rank=rank, world_size=world)
idx = dz.shard_indices(num_frames, rank, world)
K = 4
boxes = torch.zeros(len(idx), K, 9)
counts = torch.zeros(len(idx), dtype=torch.int32)
for j, f in enumerate(idx):
    boxes[j, :, 0] = f                              # tag every row with its frame id
    counts[j] = f % K + 1
b, c = dz.gather_sequence_boxes(boxes, counts, num_frames)
sg = dz.SequenceGather(num_frames, K=K, device='cpu')         # one flat buffer, ONE collective; producers write into slot views
for j0 in range(0, len(idx), 2):
    bs, cs = sg.slot(j0, min(j0 + 2, len(idx)))
    bs.copy_(boxes[j0:j0 + 2])
    cs.copy_(counts[j0:j0 + 2])
b2, c2 = sg.gather()
if rank == 0:
    with open(out, 'w') as fh:
        json.dump({'tags': b[:, 0, 0].tolist(), 'counts': c.tolist(), 'tags2': b2[:, 0, 0].tolist(), 'counts2': c2.tolist()}, fh)
dist.barrier()
dist.destroy_process_group()
