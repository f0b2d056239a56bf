"""world_size-2 gloo run of the N>1 data path (shard -> per-rank sampling -> ONE gather).
The per-rank sampler is a deterministic stand-in (the CUDA model has no CPU path); what is
under test is the sharding / gather host logic of cbgbench_b200/sharding.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cbgbench_b200 import sharding, synthetic


def _fake_sampler(sub):
    x = sub['ligand_pos'] * 2.0 + sub['ligand_element_batch'][:, None].float() * 0.0 + 1.0
    v = (sub['ligand_atom_type'] + 1) % 13
    return x, v


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        batch = synthetic.make_batch([30, 80, 10, 55, 41], [5, 9, 3, 7, 4], seed=5)
        calls = {'n': 0}
        orig = dist.all_gather

        def counting(*a, **k):
            calls['n'] += 1
            return orig(*a, **k)

        dist.all_gather = counting
        x, v, gid = sharding.sample_sharded(_fake_sampler, batch)
        dist.all_gather = orig
        torch.save({'x': x, 'v': v, 'gid': gid, 'collectives': calls['n']}, os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_sharded_sampling_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    batch = synthetic.make_batch([30, 80, 10, 55, 41], [5, 9, 3, 7, 4], seed=5)
    want_x, want_v = _fake_sampler(batch)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f'r{r}.pt'))
        assert got['collectives'] == 1                         # a single gather, nothing else
        assert torch.equal(got['gid'], batch['ligand_element_batch'])
        assert torch.equal(got['x'], want_x) and torch.equal(got['v'], want_v)
