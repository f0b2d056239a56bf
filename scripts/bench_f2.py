#!/usr/bin/env python
"""Measurement of SURVEY.md section 8 row f2 (DiffSBDD / DiffBP samplers) next to the TargetDiff path.

One GPU, workload c2 shape (64 pockets x (300 + 24) atoms, T = 1000 schedule), device-resident inputs, K timed
reverse steps per model, each bracketed by CUDA events with an L2 flush in between (same method as bench.py).
Prints one JSON line per model: ms/step, ligands/s (= B / (T * step)), launches/step, the per-kernel-family
breakdown, and the CPU port (oracle/diffusion_{sbdd,bp}.py, torch CPU) timed on ONE pocket for 2 steps.

    python scripts/bench_f2.py [--steps 20] [--warmup 3] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

T = 1000
B, N_PROT, N_LIG = 64, 300, 24


def cpu_port(model_name, threads=16):
    import torch
    import torch.nn.functional as F
    from cbgbench_b200 import synthetic
    torch.set_num_threads(min(threads, os.cpu_count() or 1))
    batch = synthetic.make_batch([N_PROT], [N_LIG], seed=2024)
    if model_name == 'diffsbdd':
        from cbgbench_b200.diffsbdd import DiffSBDDB200
        from oracle import diffusion_sbdd as O
        sd = synthetic.seeded_state_dict(DiffSBDDB200(synthetic.diffsbdd_config(num_steps=T)), seed=0)
        noise = synthetic.make_sbdd_noise(T, N_LIG, 13, seed=1)
        run = lambda n: O.sample(sd, batch, T, noise, stop_after=n)
    else:
        from cbgbench_b200.diffbp import DiffBPB200
        from oracle import diffusion_bp as O
        sd = synthetic.seeded_state_dict(DiffBPB200(synthetic.diffbp_config(num_steps=T)), seed=0)
        pn, tu = synthetic.make_bp_noise(T, N_LIG, seed=1)
        run = lambda n: O.sample(sd, batch, T, pn, tu, stop_after=n)
    run(1)
    t0 = time.perf_counter()
    run(2)
    return (time.perf_counter() - t0) / 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    import torch
    from cbgbench_b200 import _lib, synthetic
    from cbgbench_b200.targetdiff import TargetDiffB200
    from cbgbench_b200.diffsbdd import DiffSBDDB200
    from cbgbench_b200.diffbp import DiffBPB200
    torch.set_grad_enabled(False)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    batch = synthetic.make_batch([N_PROT] * B, [N_LIG] * B, seed=2024)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    t_seq = list(reversed(range(T)))
    for name, cls, cfg in (('targetdiff', TargetDiffB200, synthetic.targetdiff_config),
                           ('diffsbdd', DiffSBDDB200, synthetic.diffsbdd_config),
                           ('diffbp', DiffBPB200, synthetic.diffbp_config)):
        model = cls(cfg(num_steps=T))
        model.load_state_dict(synthetic.seeded_state_dict(model, seed=0), strict=True)
        model = model.to(dev).eval()
        torch.manual_seed(2024)
        if name == 'diffsbdd':
            state = model.begin(batch)
            step = lambda ts: model.run_steps(state, ts)
        else:
            if name == 'diffbp':
                batch_m = dict(batch)
                batch_m['ligand_atom_type'] = torch.zeros_like(batch['ligand_atom_type'])   # absorbing start
            else:
                batch_m = batch
            state = model.prepare(batch_m)
            X = torch.empty((T + 1, state['n_lig'], 3), device=dev)
            Cc = torch.empty((T + 1, state['n_lig'], 13), device=dev)
            X[T].copy_(state['x_lig'])
            Cc[T].copy_(state['c_lig'])
            step = lambda ts: model.run_steps(state, ts, X, Cc)
        step(t_seq[:args.warmup])
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        l0 = L.cbg_launch_count()
        for i in range(args.steps):
            flush.zero_()
            starts[i].record()
            step([t_seq[args.warmup + i]])
            ends[i].record()
        torch.cuda.synchronize()
        launches = (L.cbg_launch_count() - l0) / args.steps
        ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) / args.steps
        # per-kernel-family breakdown over a few extra steps
        L.cbg_profile_enable(1)
        nprof = 3
        step(t_seq[args.warmup + args.steps: args.warmup + args.steps + nprof])
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        L.cbg_profile_enable(0)
        kern = {k: round(v[0] / nprof, 4) for k, v in prof.items() if v[0] > 0}
        out = {'model': name, 'workload': f'c2 shape: {B} pockets x ({N_PROT}+{N_LIG}) atoms, T={T}', 'n_gpus': 1,
               'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms, 4),
               'ligands_per_s': round(B / (T * ms * 1e-3), 3), 'launches_per_step': launches,
               'kernel_ms_per_step': kern, 'dtype': 'f32', 'data': 'synthetic',
               'rcache': bool(state['plan'].rcache), 'l2_flush_between_steps': True}
        if not args.no_cpu and name != 'targetdiff':
            s = cpu_port(name)
            out['cpu_port'] = {'s_per_step_one_pocket': round(s, 3), 'ligands_per_s': round(1.0 / (T * s), 6),
                               'threads': min(16, os.cpu_count() or 1), 'sample': '1 pocket, 2 steps after 1 warm-up'}
        print(json.dumps(out), flush=True)
        del model, state


if __name__ == '__main__':
    main()
