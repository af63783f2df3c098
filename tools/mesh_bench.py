"""Secondary measurement: voxels/s of the 256^3 density grid (mesh extraction, BASELINE.json configs[4]), one GPU
or one slab per rank under torchrun."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anerf_b200 import mesh, parallel, synthetic
sys.argv = sys.argv[:1] + [a for a in sys.argv[1:]]
import bench as B
rank, world, local = parallel.init_distributed()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
import collections, contextlib, io
from anerf_b200.raycasters import create_raycaster
Skel = collections.namedtuple("Skel", ["joint_names", "joint_trees", "root_id"])
da = dict(skel_type=Skel(synthetic.SMPL_JOINT_NAMES, synthetic.SMPL_PARENTS, 0), near=0., far=1., n_views=1,
          joint_coords=np.tile(np.eye(3, dtype=np.float32), (1, 24, 1, 1)))
with contextlib.redirect_stdout(io.StringIO()):
    _, rk, _, _, _, _ = create_raycaster(B.make_args(), da, device=dev)
rc = rk["ray_caster"].eval()
rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
pose = synthetic.make_pose(0, 24)
kps, skts = torch.as_tensor(pose["kps"]).to(dev)[None], torch.as_tensor(pose["skts"]).to(dev)[None]
res = 255
for _ in range(2):
    g = mesh.density_grid_sharded(rc, kps, skts, 1.8, res, rank, world)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g = mesh.density_grid_sharded(rc, kps, skts, 1.8, res, rank, world)
e1.record(); torch.cuda.synchronize()
ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev)
if rank == 0:
    vox = (res + 1) ** 3
    print(json.dumps({"metric": "voxels/s, 256^3 density grid (24 joints, 8x256 trunk)", "value": vox / (ms * 1e-3), "ms": ms,
                      "n_gpus": world, "tflops_algorithmic": vox * 1.3604e6 / (ms * 1e-3) / 1e12,
                      "frac_positive": float((g > 0).float().mean())}))
if torch.distributed.is_initialized():
    torch.distributed.destroy_process_group()
