"""Quick check of the hybrid_vision training step: eager vs CUDA-graph step time, loss trajectory.  python tools/train_step_check.py [batch]
(under torchrun: DDP)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import hvs_b200
from hvs_b200 import harness
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    dist.init_process_group("nccl", device_id=dev)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["eager", "graph"]
for mode in modes:
    model = harness.build_model(dev, seed=0)
    r = harness.training_ddp(model, dev, world, rank, batch, 640, steps=4, warmup=2, use_graph=mode == "graph")
    if rank == 0:
        print(mode, json.dumps({k: r[k] for k in ("ms_per_step", "loss_first", "loss_last", "finite", "hvs_launches_per_step", "peak_mem_gb", "cuda_graph")}), flush=True)
    del model
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
