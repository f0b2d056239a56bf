#!/usr/bin/env python
"""Row f3 measurement: building one pocket's sampling batches (sample.py:177-183: num_samples evaluations of the
transform list + collate) on the GPU vs the CPU restatement of the reference's per-sample transform passes (oracle,
torch CPU - the reference's own formulation: pdist / sort / median, numpy choice, per-sample tensors, concatenation).
Prints one JSON line.   python scripts/bench_f3.py [--atoms 400] [--samples 200]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--atoms', type=int, default=400)
    ap.add_argument('--samples', type=int, default=200)      # configs/denovo/test/targetdiff.yml: num_samples 200
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    from cbgbench_b200 import _lib
    from cbgbench_b200.batch_builder import DeviceBatchBuilder, SizePrior
    from oracle import batch_builder as OB
    G = np.load(os.path.join(ROOT, 'tests', 'golden', 'batch_builder.npz'))
    ptr = G['table/bin_ptr']
    tab = {'bounds': list(G['table/bounds']),
           'bins': [(list(G['table/values'][ptr[b]:ptr[b + 1]]), list(G['table/probs'][ptr[b]:ptr[b + 1]])) for b in range(len(ptr) - 1)]}
    prior = SizePrior.from_table(tab)
    rs = np.random.RandomState(0)
    n = args.atoms
    pocket = {'pos': torch.from_numpy((rs.normal(0, 5.2, size=(n, 3)) + 17.0).astype(np.float32)),
              'element': torch.from_numpy(rs.choice([1, 6, 7, 8, 16], size=n)), 'is_backbone': torch.from_numpy(rs.randint(0, 2, size=n).astype(bool)),
              'atom_to_aa_type': torch.from_numpy(rs.randint(0, 20, size=n))}
    dev = torch.device('cuda:0')
    b = DeviceBatchBuilder(prior, recipe='denovo', type_dist='uniform', pos_dist='gaussian')
    dpocket = {k: v.to(dev) for k, v in pocket.items()}
    for _ in range(3):
        out = b.build([dpocket], args.samples, device=dev)
    torch.cuda.synchronize()
    L = _lib.lib()
    l0 = L.cbg_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        out = b.build([dpocket], args.samples, device=dev)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / args.reps * 1e3
    launches = (L.cbg_launch_count() - l0) / args.reps
    # CPU: the reference's formulation, one transform pass per sample + collate
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    t0 = time.perf_counter()
    feat, aa = OB.featurize_protein(pocket['element'], pocket['atom_to_aa_type'], pocket['is_backbone'])
    samples = []
    for s in range(args.samples):
        feat, aa = OB.featurize_protein(pocket['element'], pocket['atom_to_aa_type'], pocket['is_backbone'])   # dataset[i] re-runs it
        u = rs.random_sample()
        p, centre, x, t = OB.denovo_sample(pocket['pos'], tab, u, 'uniform', torch.rand(80, 13), 'gaussian', torch.randn(80, 3), 13)
        samples.append({'protein_pos': p, 'centre': centre, 'ligand_pos': x, 'ligand_atom_type': t})
    OB.collate(samples, feat, aa)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({'row': 'f3', 'workload': f'1 pocket x {n} atoms, {args.samples} samples (de-novo recipe)',
                      'gpu_ms_per_batch': gpu_ms, 'gpu_launches_per_batch': launches,
                      'cpu_port_ms_per_batch': cpu_ms, 'cpu_threads': torch.get_num_threads(),
                      'speedup': cpu_ms / gpu_ms, 'ligand_atoms': int(out['ligand_pos'].shape[0]),
                      'note': 'GPU time is wall clock incl. the torch RNG draws and the one size read-back; the CPU side is the '
                              'oracle restatement (the real reference additionally pays PyG Data/Batch object overhead)'}))


if __name__ == '__main__':
    main()
