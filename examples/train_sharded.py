"""Training over a destination-partitioned graph, one process per GPU (the scaling design of BASELINE.json's north
star, SURVEY 8e, with the gradients of SURVEY 8f rank 3):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 examples/train_sharded.py

Every rank holds the whole edge list and the features of all nodes here (a pre-sharded graph would go through
`ShardedForward.forward_ids`), owns the rows of one destination range, computes a loss over those rows, and sums its
share of every parameter gradient with the other ranks' before the optimiser step - the model parameters stay
replicated and identical.  Runs on one GPU too (world size 1).  Needs the f16 engine at hidden_dim 128.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-hypernetwork-forge_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from graph_hypernetwork_forge import HyperGNN  # noqa: E402
from graph_hypernetwork_forge.distributed import ShardedForward  # noqa: E402


def main(steps: int = 10, verbose: bool = True):
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank()
    # the same synthetic graph and the same initial parameters on every rank
    N, E, R, F, d = 40_000, 300_000, 11, 32, 128
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(N, F, generator=g, device=dev)
    ei = torch.randint(0, N, (2, E), generator=g, device=dev)
    names = [f"relation number {r}" for r in range(R)]
    texts = [names[int(r)] for r in torch.randint(0, R, (E,), generator=g, device=dev).tolist()]
    target = torch.randn(N, d, generator=g, device=dev)
    torch.manual_seed(0)
    model = HyperGNN(text_dim=64, node_feat_dim=F, hidden_dim=d, num_layers=2, precision="f16").to(dev).train()
    sharded = ShardedForward(model, N)                   # equal node ranges; see plan_partition_by_edges for skew
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    losses = []
    for step in range(steps):
        opt.zero_grad()
        out = sharded.forward(x, ei, texts)              # the rank's own rows [hi - lo, d]
        loss = ((out - target[sharded.lo:sharded.hi]) ** 2).sum() / (N * d)
        loss.backward()
        sharded.allreduce_gradients()                    # this rank's share + everybody else's
        opt.step()
        total = loss.detach().clone()
        dist.all_reduce(total)                           # the loss over all rows, for the log
        losses.append(float(total))
        if verbose and rank == 0 and (step + 1) % 5 == 0:
            print(f"step {step + 1:3d}  loss {losses[-1]:.5f}", flush=True)
    # replicas stayed identical: same parameters on every rank after the same updates
    check = torch.cat([p.detach().flatten()[:64] for p in model.parameters()])
    lo, hi = check.clone(), check.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "parameters diverged between ranks"
    return losses


if __name__ == "__main__":
    out = main()
    assert out[-1] < out[0], out
    dist.destroy_process_group()
