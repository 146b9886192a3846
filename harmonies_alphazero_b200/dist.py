"""Multi-GPU self-play: one process per GPU, games sharded by global game id, and the only
two exchanges the path has (both new: the reference ships weights and trajectories through
pickled multiprocessing tasks, trainer.py:76-117):

* ``broadcast_weights`` — after each training step the trainer rank broadcasts ONE flat
  buffer with all parameters and BatchNorm statistics (2.45 M values for the default net);
* ``gather_trajectories`` — packed examples (128 B state + int16 visit counts + z) are
  gathered to the replay-buffer rank and re-encoded there with hz_encode, ~0.4 KB per example
  instead of the reference's 6 KB fp32 tuple.

Works on any torch.distributed backend: NCCL over NVLink/NVSwitch on the GPU box, gloo in the
CPU tests.  Self-play itself needs no collective: games are independent.
"""

import torch
import torch.distributed as dist

from .selfplay import Trajectories


def shard_games(num_games, rank, world):
    """Contiguous game-id range of ``rank``: (first_id, count).  Ranges tile [0, num_games)."""
    base, extra = divmod(num_games, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def _flat_tensors(model):
    return [p.data for p in model.parameters()] + [b.data for b in model.buffers() if b.dtype.is_floating_point]


def broadcast_weights(model, src=0, group=None, dtype=None):
    """Broadcast every parameter and floating-point buffer of ``model`` from ``src`` in one
    flat message.  ``dtype`` (e.g. torch.bfloat16) halves the bytes when the receivers only
    run bf16 inference.  Returns the number of bytes sent per rank.

    The flat wire buffer and its per-tensor views are built once per (model, dtype) and reused:
    a call is one fused pack (``torch._foreach_copy_``), one collective and one fused unpack,
    not two small kernels per tensor."""
    tensors = _flat_tensors(model)
    if not tensors:
        return 0
    dev = tensors[0].device
    wire = dtype or torch.float32
    cache = model.__dict__.setdefault("_hz_wire_cache", {})
    key = (wire, dev, tuple(t.numel() for t in tensors))
    if key not in cache:
        flat = torch.empty(sum(t.numel() for t in tensors), dtype=wire, device=dev)
        views, off = [], 0
        for t in tensors:
            views.append(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
        cache.clear()
        cache[key] = (flat, views)
    flat, views = cache[key]
    if dist.get_rank(group) == src:
        torch._foreach_copy_(views, tensors)
    dist.broadcast(flat, src=src, group=group)
    if dist.get_rank(group) != src:
        torch._foreach_copy_(tensors, views)
    return flat.numel() * flat.element_size()


def gather_trajectories(traj, dst=None, group=None):
    """Collect every rank's Trajectories.  dst=None: all ranks get everything (all_gather);
    otherwise only ``dst`` does and the others get an empty set.  Rows keep their global
    game ids, so the result is independent of the number of ranks up to row order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = traj.states.device
    # example counts of all ranks: one collective into one tensor and ONE host read (the sizes are needed on the host)
    n = torch.tensor([len(traj)], dtype=torch.int64, device=dev)
    all_n = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_n, n, group=group)
    counts = all_n.tolist()
    m = max(counts) if counts else 0
    # one packed row per example: state (32 x i32) | visits (143 x i16 -> 72 x i32) | game_id(2) | z | move
    row = torch.zeros((m, 32 + 72 + 1 + 2 + 1), dtype=torch.int32, device=dev)
    k = len(traj)
    if k:
        row[:k, :32] = traj.states
        v = torch.zeros((k, 144), dtype=torch.int16, device=dev)
        v[:, :143] = traj.visits
        row[:k, 32:104] = v.view(torch.int32)
        row[:k, 104:106] = traj.game_id.contiguous().view(torch.int32).view(k, 2)
        row[:k, 106] = traj.z.contiguous().view(torch.int32)
        row[:k, 107] = traj.move_no.to(torch.int32)
    # one all-gather into one buffer for both modes: over NVSwitch the ring all-gather of the padded rows (1.4 ms for
    # 8 x 13.7 MB) is faster than NCCL's send/recv gather to one root (3.0 ms measured), so ranks other than dst simply
    # drop what they received
    buf = torch.empty((world * m, row.shape[1]), dtype=torch.int32, device=dev)     # concatenated form (gloo accepts only this one)
    dist.all_gather_into_tensor(buf, row, group=group)
    rows = list(buf.view(world, m, row.shape[1]).unbind(0))
    if dst is not None and rank != dst:
        rows, counts = [], []
    parts = [r[:c] for r, c in zip(rows, counts) if c]
    allr = torch.cat(parts) if parts else row[:0]
    t = allr.shape[0]
    stats = dict(traj.stats)
    stats["bytes_per_example"] = row.shape[1] * 4
    stats["gathered_examples"] = t
    return Trajectories(
        states=allr[:, :32].clone(),
        visits=allr[:, 32:104].clone().view(torch.int16)[:, :143].clone(),
        z=allr[:, 106].clone().view(torch.float32),
        game_id=allr[:, 104:106].clone().view(torch.int64).view(t),
        move_no=allr[:, 107].clone(),
        stats=stats,
    )


def sharded_self_play(net, cfg, num_games, device, group=None):
    """Each rank plays its shard of ``num_games`` with its own BatchedSelfPlay and returns
    its local Trajectories (call gather_trajectories to collect them)."""
    import dataclasses

    from .selfplay import BatchedSelfPlay

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = shard_games(num_games, rank, world)
    local = dataclasses.replace(cfg, first_game_id=cfg.first_game_id + first)
    return BatchedSelfPlay(net, local, device=device).play(count)
