"""2-rank check of the bucketed, overlapped gradient exchange (run under torchrun): after 3 graphed steps the weights
are identical on every rank, and bit-identical to the run with a single all-reduce (STARCOP_NO_GRAD_BUCKETS=1)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from starcop_b200 import parallel, synthetic
from starcop_b200.model_setup import get_model
from starcop_b200.settings import default_settings
rank, lr, world = parallel.init_distributed("nccl")
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
sums = {}
for mode in ("bucketed", "single"):
    torch.manual_seed(0)
    m = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(dev).train()
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic.hyperstarcop_batch(4, size=256, seed=10 + rank).items()}
    gs = parallel.GradSync(world, bucketed=(mode == "bucketed"))
    step = m.make_graphed_train_step(b, grad_sync=gs)
    for _ in range(3):
        loss = step(b)
    torch.cuda.synchronize()
    h = hashlib.sha1(m.network.flat_params.cpu().numpy().tobytes()).hexdigest()[:16]
    gathered = [None] * world
    dist.all_gather_object(gathered, h)
    assert len(set(gathered)) == 1, (mode, gathered)
    sums[mode] = h
    del step
    torch.cuda.synchronize()
if rank == 0:
    print("weights after 3 steps:", sums, "identical across ranks; bucketed == single:", sums["bucketed"] == sums["single"])
assert sums["bucketed"] == sums["single"]
dist.barrier()
dist.destroy_process_group()
