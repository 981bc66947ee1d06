"""Multi-GPU plumbing for the self-play path: one process per GPU, torch.distributed (NCCL over NVLink).

The search itself never communicates: every game / tree is independent, so games are sharded
across ranks with no data-path collective.  Collectives happen once per generation only:
  broadcast_weights  -- new network weights from the training rank (replaces pickling a deepcopy(net) into
                        every handler process, examplegenerator.py:121)
  gather_records     -- variable-length training-record buffers to every rank (replaces Pool.map_async
                        result pickling, examplegenerator.py:130-136,152)
On CPU tensors (gloo) the same code paths run for the world_size-2 tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(n_items, rank=None, world=None):
    """Contiguous share [lo, hi) of n_items for this rank."""
    if rank is None:
        rank, world = rank_world()
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _comm_device(device):
    backend = dist.get_backend()
    return torch.device(device) if backend == "nccl" else torch.device("cpu")


@torch.no_grad()
def broadcast_weights(net, src=0, device=None):
    """In-place broadcast of every parameter and buffer of `net` from rank `src` (one flat fp32 bucket: the nets
    are 0.9-4 MB, so a single collective is launch-latency bound, not bandwidth bound)."""
    rank, world = rank_world()
    if world == 1:
        return net
    tensors = [t for t in list(net.parameters()) + list(net.buffers())]
    dev = _comm_device(device if device is not None else tensors[0].device)
    flat = torch.cat([t.detach().reshape(-1).to(dev, torch.float32) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].reshape(t.shape).to(t.device, t.dtype))
        off += n
    return net


def gather_records(records, device="cpu"):
    """all_gather of variable-length structured record arrays (two-phase: lengths, then padded payload)."""
    rank, world = rank_world()
    if world == 1:
        return records
    dev = _comm_device(device)
    dtype = records.dtype
    raw = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(-1).copy()).to(dev)
    n = torch.tensor([raw.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros((cap,), dtype=torch.uint8, device=dev)
    padded[:raw.numel()] = raw
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    parts = [bufs[r][:sizes[r]].cpu().numpy().view(dtype) for r in range(world)]
    out = np.concatenate(parts)
    # disambiguate trees of different ranks: rank r's tree ids are offset by r * 2^20
    off = 0
    for r in range(world):
        k = sizes[r] // dtype.itemsize
        out["tree"][off:off + k] += r * (1 << 20)
        off += k
    return out


@torch.no_grad()
def allreduce_gradients(net, device=None):
    """Data-parallel training step (Trainer(ddp=True)): average the gradients of `net` over the ranks with ONE all-reduce of
    a flat fp32 bucket (the nets have 0.2-1 M parameters: launch-latency bound, NVLS in the switch when NCCL enables it)."""
    rank, world = rank_world()
    if world == 1:
        return
    params = [p for p in net.parameters() if p.grad is not None]
    if not params:
        return
    dev = _comm_device(device if device is not None else params[0].device)
    flat = torch.cat([p.grad.reshape(-1).to(dev, torch.float32) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].reshape(p.shape).to(p.grad.device, p.grad.dtype))
        off += n
