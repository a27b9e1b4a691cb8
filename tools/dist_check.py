"""torchrun --nproc-per-node 2 tools/dist_check.py : sharded fit (+ one NCCL all-reduce, factor-sharded eigh) == single-GPU fit."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.utils.data import DataLoader, TensorDataset
from laplace_b200 import models
from laplace_b200.distributed import fit_distributed
from laplace_b200.posterior import B200Laplace

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = models.make("resnet18", width=16).cuda()
torch.manual_seed(1)
X, y = torch.randn(512, 3, 32, 32).cuda(), torch.randint(10, (512,)).cuda()
loader = DataLoader(TensorDataset(X, y), batch_size=64)
la = fit_distributed(B200Laplace(model, "classification", "all", "kron"), loader)
f_mu, f_var = la.glm_predictive_distribution(X[:8])
if rank == 0:
    ref = B200Laplace(model, "classification", "all", "kron").fit(loader)
    worst = max(float((a - b).norm() / b.norm()) for Fa, Fb in zip(la.H_facs.kfacs, ref.H_facs.kfacs) for a, b in zip(Fa, Fb))
    fv = ref.glm_predictive_distribution(X[:8])[1]
    print(f"dist_check world={world}: factors rel-fro vs single GPU {worst:.2e}; loss {float(la.loss):.4f} vs {float(ref.loss):.4f}; "
          f"f_var rel {float((f_var - fv).abs().max() / fv.abs().max()):.2e}")
    assert worst < 1e-5 and abs(float(la.loss) - float(ref.loss)) < 1e-2 * abs(float(ref.loss))
dist.barrier()
dist.destroy_process_group()
