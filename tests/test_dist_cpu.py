"""CPU, gloo, world_size 2: the multi-rank path of dist.py — game sharding, flat weight
broadcast, packed trajectory gather (the collectives run over NCCL on the GPU box)."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harmonies_alphazero_b200.dist import broadcast_weights, gather_trajectories, shard_games
from harmonies_alphazero_b200.net import AlphaZeroNet, TEST_MODEL_CONFIG
from harmonies_alphazero_b200.selfplay import Trajectories


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_traj(rank, n):
    g = torch.Generator().manual_seed(100 + rank)
    return Trajectories(
        states=torch.randint(-2**31, 2**31 - 1, (n, 32), dtype=torch.int32, generator=g),
        visits=torch.randint(0, 400, (n, 143), dtype=torch.int16, generator=g),
        z=torch.randint(-1, 2, (n,), generator=g).float(),
        game_id=torch.arange(n, dtype=torch.int64) + (1 << 33) * rank,
        move_no=torch.arange(n, dtype=torch.int32) % 70,
        stats={"sims": n},
    )


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weights: rank 0's values must arrive everywhere, bit-exact in fp32
        torch.manual_seed(rank)
        m = AlphaZeroNet.from_config(TEST_MODEL_CONFIG)
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.uniform_(0.1, 2.0)
        nbytes = broadcast_weights(m, src=0)
        torch.manual_seed(0)
        ref = AlphaZeroNet.from_config(TEST_MODEL_CONFIG)
        ok_w = all(torch.equal(a, b) for a, b in zip(m.parameters(), ref.parameters())) if rank else True
        # a second broadcast after a training step on rank 0 carries the NEW values (the cached
        # wire buffer is refilled on every call)
        if rank == 0:
            with torch.no_grad():
                for p_ in m.parameters():
                    p_.add_(1.0)
        broadcast_weights(m, src=0)
        ok_w = ok_w and all(torch.equal(a, b + 1.0) for a, b in zip(m.parameters(), ref.parameters()))
        # bf16 wire format halves the bytes
        nbytes16 = broadcast_weights(m, src=0, dtype=torch.bfloat16)
        # trajectories: ragged counts (rank 1 has more, one rank may have none)
        mine = _fake_traj(rank, [5, 9][rank])
        allt = gather_trajectories(mine)
        want = [_fake_traj(r, [5, 9][r]) for r in range(world)]
        ok_t = (
            torch.equal(allt.states, torch.cat([w.states for w in want]))
            and torch.equal(allt.visits, torch.cat([w.visits for w in want]))
            and torch.equal(allt.z, torch.cat([w.z for w in want]))
            and torch.equal(allt.game_id, torch.cat([w.game_id for w in want]))
            and torch.equal(allt.move_no, torch.cat([w.move_no for w in want]))
        )
        only0 = gather_trajectories(mine, dst=0)
        ok_d = len(only0) == (14 if rank == 0 else 0)
        empty = gather_trajectories(_fake_traj(rank, 0 if rank == 0 else 3))
        ok_e = len(empty) == 3
        out[rank] = (ok_w, ok_t, ok_d, ok_e, nbytes, nbytes16)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_broadcast_and_gather():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()      # never fork a process that already runs torch threads
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        ok_w, ok_t, ok_d, ok_e, nb, nb16 = out[r]
        assert ok_w and ok_t and ok_d and ok_e, (r, out[r])
        assert nb == 2 * nb16 and nb > 4 * 50985      # 50,985 parameters + BN statistics


def test_shard_games_tiles_the_range():
    for n, w in [(32768, 8), (10, 4), (3, 8), (0, 2), (4096, 1)]:
        spans = [shard_games(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (f0, c0), (f1, _) in zip(spans[:-1], spans[1:]):
            assert f1 == f0 + c0
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
