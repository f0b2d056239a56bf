"""Sampling driver: the B200 counterpart of the reference's ``sample.py`` loop.

Mirrors /root/reference sample.py:159-241 for the part that belongs to the hot path: build mini-batches
of pocket+ligand graphs, ``model.sample(batch)``, take ``traj[0]`` (the state the reference consumes,
sample.py:194-201), split it per pocket (``split_batch_into_samples``, sample.py:16-32) and collect
``{pos, v}`` per ligand.  What stays outside: LMDB/PDB parsing and RDKit/OpenBabel reconstruction (CPU
chemistry tooling, out of scope) - pockets are synthetic (``cbgbench_b200.synthetic``)
or come from a ``torch.save``d list of batch dicts with the reference's keys.

Multi-GPU (torchrun, one rank per GPU): every mini-batch's pockets are partitioned over the ranks by atom
count, each rank samples its share with no communication and ONE all-gather returns the final
coordinates / types (``cbgbench_b200.sharding``).

    python -m cbgbench_b200.sample_driver --pockets 16 --batch-size 16 --out out.pt
    python -m cbgbench_b200.sample_driver --model diffsbdd --pockets 16      (row f2: also diffbp)
    torchrun --nproc-per-node 8 -m cbgbench_b200.sample_driver --pockets 512 --batch-size 512
"""
import argparse
import os
import time

import torch

from . import sharding, synthetic
from .targetdiff import TargetDiffB200
from .diffsbdd import DiffSBDDB200
from .diffbp import DiffBPB200

MODELS = {'targetdiff': (TargetDiffB200, synthetic.targetdiff_config),
          'diffsbdd': (DiffSBDDB200, synthetic.diffsbdd_config),
          'diffbp': (DiffBPB200, synthetic.diffbp_config)}


def split_batch_into_samples(x, v, graph_id, n_graphs):
    """Per-pocket results of one mini-batch (sample.py:16-32): list of dicts {pos [n,3], v [n]}."""
    out = []
    for g in range(n_graphs):
        m = graph_id == g
        out.append({'pos': x[m].cpu(), 'v': v[m].cpu()})
    return out


def final_state(model, sub_batch, traj_key=0):
    """model.sample on a (sub-)batch -> (x, v) of traj[traj_key] on the model's device."""
    traj = model.sample(sub_batch, traj_mode='final')
    x, c, _ = traj[traj_key]
    dev = next(model.parameters()).device
    return x.to(dev), c.argmax(-1).to(dev)


def run(args):
    distributed = int(os.environ.get('WORLD_SIZE', '1')) > 1
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if distributed and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl', device_id=dev)
    cls, make_cfg = MODELS[args.model]                                                     # get_model(cfg), sample.py:155
    model = cls(make_cfg(num_steps=args.steps, num_layers=args.layers))
    if args.ckpt:
        ckpt = torch.load(args.ckpt, map_location='cpu')
        model.load_state_dict(ckpt['model'] if 'model' in ckpt else ckpt, strict=True)     # sample.py:153-157
    else:
        model.load_state_dict(synthetic.seeded_state_dict(model, seed=0), strict=True)
    model = model.to(dev).eval()
    torch.manual_seed(args.seed + rank)                                                    # sample.py:106,131-133

    if args.batches:
        batches = torch.load(args.batches)
    else:
        batches = []
        for b0 in range(0, args.pockets, args.batch_size):
            nb = min(args.batch_size, args.pockets - b0)
            b = synthetic.make_batch([args.n_prot] * nb, [args.n_lig] * nb, seed=args.seed + b0, gen_mode=args.gen_mode)
            if args.model == 'diffbp':        # configs/denovo/test/diffbp.yml:19-21 assign_atomtype: absorbing
                b['ligand_atom_type'] = torch.zeros_like(b['ligand_atom_type'])
            batches.append(b)
    results, t0 = [], time.time()
    for batch in batches:
        n_graphs = int(batch['ligand_element_batch'].max()) + 1
        if distributed:
            x, v, gid = sharding.sample_sharded(lambda sub: final_state(model, sub), batch)
        else:
            x, v = final_state(model, batch)
            gid = batch['ligand_element_batch'].to(x.device)
        # translate back like sample.py:198-201 (synthetic pockets are already centred: translation = 0)
        tr = batch.get('protein_translation')
        if tr is not None:
            x = x + tr.to(x.device)[gid]
        results.extend(split_batch_into_samples(x, v, gid, n_graphs))
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        if args.out:
            torch.save(results, args.out)
        print(f'sampled {len(results)} ligands in {dt:.2f} s ({len(results) / dt:.2f} ligands/s, '
              f'{args.steps} steps, world {int(os.environ.get("WORLD_SIZE", "1"))})')
    if distributed:
        torch.distributed.barrier()
    return results


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--model', default='targetdiff', choices=sorted(MODELS))
    ap.add_argument('--pockets', type=int, default=16)
    ap.add_argument('--batch-size', type=int, default=16)          # sample.py:108
    ap.add_argument('--n-prot', type=int, default=300)
    ap.add_argument('--n-lig', type=int, default=24)
    ap.add_argument('--gen-mode', default='denovo', choices=['denovo', 'partial'])
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--layers', type=int, default=9)
    ap.add_argument('--seed', type=int, default=2024)               # sample.py:106
    ap.add_argument('--ckpt', default=None, help='reference checkpoint ({"model": state_dict, ...})')
    ap.add_argument('--batches', default=None, help='torch.save()d list of batch dicts with the reference keys')
    ap.add_argument('--out', default=None)
    return run(ap.parse_args(argv))


if __name__ == '__main__':
    main()
