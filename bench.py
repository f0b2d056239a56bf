#!/usr/bin/env python
"""Benchmark of the diffusion-sampling hot path (BASELINE.json metric: ligands/sec sampled,
1000-step denoise, batch 64 pockets of ~300 protein + 24 ligand atoms per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1: one rank per GPU, weak scaling

A "step" is ONE denoise step of the whole batch: ligand embedding -> device kNN -> edge gate ->
9 x (node GEMMs, fused X2H, fused H2X) -> classifier -> reverse diffusion step (+ the two torch
RNG draws).  value = ligands/s = (64 x N) / (T x seconds per step), T = 1000.
Timing: W >= 3 warm-up steps, then K steps each bracketed by CUDA events on the launching
stream with an L2 flush (256 MiB write) between timed steps, barrier + synchronize around the
region, max over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_STEPS = 1000
WORKLOADS = {
    # name: (graphs per GPU, protein atoms, ligand atoms, gen_mode, encoder overrides, description)
    'c2': (64, 300, 24, 'denovo', {}, 'de novo denoise, batch 64 pockets x (300 protein + 24 ligand atoms), 1000 steps, fp32'),
    'c3': (128, 300, 24, 'partial', {'cutoff_mode': 'radius', 'r_max': 10.0},
           'linker task, batch 128, radius graph r=10 A (cap 32), 1000 steps'),
    'c1': (1, 200, 24, 'denovo', {}, 'de novo, 1 pocket 200+24 atoms (plumbing case)'),
    # ragged pockets 100..800 protein atoms (mean 450), 24 ligand atoms, context fixed: 32 pockets per GPU (256 over 8)
    'c5': (32, None, 24, 'partial', {}, 'scaffold task, ragged pockets 100-800 atoms, 32 pockets per GPU (256 over 8 GPUs)'),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--e2e-steps', type=int, default=T_STEPS, help='denoise steps of the end-to-end sample() call')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-steps', type=int, default=3)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms (NVML) while running."""
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost'}

    def __init__(self, torch_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid).encode() if not uuid.startswith('GPU-') else uuid.encode())
            except Exception:
                vis = os.environ.get('CUDA_VISIBLE_DEVICES')
                idx = int(vis.split(',')[torch_index]) if vis and vis.split(',')[torch_index].isdigit() else torch_index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:   # NVML missing: report nulls rather than fail the bench
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                fn = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(fn(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': 0}
        return {'sm_mhz': statistics.median(self.samples), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


def workload_batch(name, rank, n_graphs=None):
    from cbgbench_b200 import synthetic
    import numpy as np
    B, n_prot, n_lig, gen_mode, enc, _ = WORKLOADS[name]
    B = n_graphs or B
    if n_prot is None:
        sizes = [int(v) for v in np.random.RandomState(77 + rank).randint(100, 801, size=B)]
    else:
        sizes = [n_prot] * B
    return synthetic.make_batch(sizes, [n_lig] * B, seed=2024 + rank, gen_mode=gen_mode), enc


# ---------------------------------------------------------------------------------------------
def cpu_reference_steps(workload, n_graphs, steps, warmup, threads):
    """Time `steps` denoise steps of the oracle port (torch CPU, fp32, as-written reference
    formulation) on n_graphs graphs of the workload shape.  Returns seconds per step."""
    import torch
    from cbgbench_b200 import synthetic
    from cbgbench_b200.targetdiff import TargetDiffB200
    from oracle import diffusion as OD
    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    batch, enc = workload_batch(workload, 0, n_graphs)
    model = TargetDiffB200(synthetic.targetdiff_config(num_steps=T_STEPS, **enc))
    sd = synthetic.seeded_state_dict(model, seed=0)
    import torch.nn.functional as F
    x = batch['ligand_pos'].float()
    c = F.one_hot(batch['ligand_atom_type'], 13).float()
    gen = batch.get('ligand_gen_flag', batch['ligand_lig_flag'])
    kw = dict(k=32, cutoff_mode=enc.get('cutoff_mode', 'knn'), r_max=enc.get('r_max', 10.0))
    times = []
    t_idx = T_STEPS - 1
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        x0, logits = OD.denoise_once(sd, batch, x, c, **kw)
        x = OD.pos_reverse_step(sd, x0, x, t_idx, gen, torch.randn_like(x))
        c, _ = OD.type_reverse_step(sd, logits, c, t_idx, gen, torch.rand_like(c), 13)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        t_idx -= 1
    return sum(times) / len(times)


def pick_cpu_sample(workload, budget_s):
    """Choose the torch thread count that runs the reference path fastest on this host (all cores is often
    slower than a moderate count for these op sizes) and the largest graph count (<= the workload's) whose
    step fits the time budget.  Returns (n_graphs, threads, tried)."""
    B = WORKLOADS[workload][0]
    ncpu = os.cpu_count() or 1
    # all-cores is pathologically slow for these op sizes on many-core hosts (measured 85 s/step for ONE pocket
    # with 128 threads vs 0.07 s with 16), so the probe is capped at 64 threads to keep the run short
    cands = sorted({c for c in (min(ncpu, 64), 32, 16, 8) if c <= ncpu}, reverse=True)
    tried = {}
    for c in cands:
        tried[c] = cpu_reference_steps(workload, 1, 1, 1, c)
    threads = min(tried, key=tried.get)
    n = max(1, min(B, int(budget_s / max(tried[threads], 1e-3))))
    return n, threads, {k: round(v, 3) for k, v in tried.items()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: torch CPU
    fp32, as-written formulation) on this box's host cores, same metric/config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    B, n_prot, n_lig, _, enc, desc = WORKLOADS[args.workload]
    total_steps = args.steps + args.warmup
    n_graphs, threads, tried = pick_cpu_sample(args.workload, budget_s=max(0.5, 120.0 / max(total_steps, 1)))
    sec = cpu_reference_steps(args.workload, n_graphs, args.steps, args.warmup, threads)
    value = n_graphs / (T_STEPS * sec)
    sample = (f'{n_graphs} of {B} pockets ({n_prot}+{n_lig} atoms each), {args.steps} denoise steps after {args.warmup} '
              f'warm-up, ligands/s = pockets / (1000 x s/step); {threads} torch threads of {os.cpu_count()} host cores '
              f'(fastest of s/step for 1 pocket: {tried})')
    line = {
        'impl': 'reference', 'metric': 'ligands/sec sampled (1000-step denoise, batch 64)', 'value': value,
        'unit': 'ligands/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': f'{args.workload}: {desc}', 'graphs_timed': n_graphs,
                                        'kind': 'oracle port of the reference CPU path (torch CPU, all host threads)'},
        'cpu_baseline': {'value': value, 'unit': 'ligands/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'ligands/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from cbgbench_b200 import _lib, synthetic
    from cbgbench_b200.targetdiff import TargetDiffB200

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)'
    if args.warmup < 3:
        raise SystemExit('--warmup must be >= 3')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    torch.set_grad_enabled(False)
    L = _lib.lib()

    B, n_prot, n_lig, gen_mode, enc, desc = WORKLOADS[args.workload]
    batch, _ = workload_batch(args.workload, rank)
    model = TargetDiffB200(synthetic.targetdiff_config(num_steps=T_STEPS, **enc))
    model.load_state_dict(synthetic.seeded_state_dict(model, seed=0), strict=True)
    model = model.to(dev).eval()
    torch.manual_seed(2024 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps (inputs already in HBM) ------------------------------------------
    state = model.prepare(batch)
    n_lig_tot, K = state['n_lig'], model.num_classes
    N = state['n_nodes']
    need = args.warmup + args.steps + args.profile_steps
    assert need < T_STEPS
    X = torch.empty((T_STEPS + 1, n_lig_tot, 3), device=dev)
    Cc = torch.empty((T_STEPS + 1, n_lig_tot, K), device=dev)
    X[T_STEPS].copy_(state['x_lig'])
    Cc[T_STEPS].copy_(state['c_lig'])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    t_seq = list(reversed(range(T_STEPS)))
    model.run_steps(state, t_seq[:args.warmup], X, Cc)
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = L.cbg_launch_count()
    with ClockSampler(local_rank) as clocks:
        for i in range(args.steps):
            flush.zero_()                                  # L2 flush between timed steps (outside the events)
            starts[i].record()
            model.run_steps(state, [t_seq[args.warmup + i]], X, Cc)
            ends[i].record()
        barrier()
    gpu_launches = L.cbg_launch_count() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    ms_per_step = sum(step_ms) / len(step_ms)
    t = torch.tensor([ms_per_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())
    value = (B * world) / (T_STEPS * ms_per_step * 1e-3)

    # ---- per-kernel CUDA-event profile of a few more steps (roofline of the dominant kernel) ------
    prof = None
    if args.profile_steps > 0:
        L.cbg_profile_enable(1)
        p0 = args.warmup + args.steps
        model.run_steps(state, t_seq[p0:p0 + args.profile_steps], X, Cc)
        prof = _lib.profile_collect()
        L.cbg_profile_enable(0)
    peaks, peak_kind = measured_peaks()
    roofline, kernels = None, None
    if prof:
        kernels = {k: {'ms_per_step': v[0] / args.profile_steps, 'launches_per_step': v[1] / args.profile_steps}
                   for k, v in prof.items() if v[1]}
        dom = max(('x2h_k', 'x2h_v'), key=lambda k: prof[k][0])
        dom_ms = prof[dom][0] / prof[dom][1]
        # ALGORITHMIC bytes per launch of the fused X2H kernels, per node:
        #   x2h_k: Pj_k + Pi_k + q rows (3 x 512 B) + nbr (128) + e_w (128) + x (16) read, w (2048) written
        #   x2h_v: Pj_v + Pi_v (2 x 512) + nbr + e_w + x (272) + w (2048) + h (512) read, h (512) written
        #   + the node's R-cache block (32 slots x 512 B) streamed for static (non-generated) nodes
        # rows one launch processes: with receptive-field pruning layer l only updates the nodes that can still reach a
        # sampled atom, so the per-launch average over the layers is what the measured launch time corresponds to
        n_layers = state['plan'].num_layers
        rows = float(N)
        if model.use_prune:
            import ctypes
            cnt = (ctypes.c_int32 * (n_layers + 1))()
            _lib.check(L.cbg_sample_prune_counts_host(ctypes.byref(state['plan']), cnt, _lib.stream_ptr(dev)))
            rows = sum(cnt[l + 1] for l in range(n_layers)) / n_layers
        rows_static = max(rows - state['plan'].n_gen, 0.0)       # generated atoms are in every layer's list
        rc_bytes = 32 * 512 * rows_static if model.use_rcache else 0
        per_node = {'x2h_k': 3 * 512 + 272 + 2048, 'x2h_v': 2 * 512 + 272 + 2048 + 1024}[dom]
        alg_bytes = per_node * rows + rc_bytes
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        # useful fp32 work of the same launch (query-folded / aggregated second Linears; the RBF mat-vec only
        # for the edges that are not served from the R-cache is NOT counted: lower bound of useful FLOPs)
        flops_node = {'x2h_k': 2 * (128 * 128 + 32 * 128 * 16), 'x2h_v': 2 * (32 * 128 * 16 + 128 * 128)}[dom]
        sm_max = (clocks.summary()['sm_max_mhz'] or peaks.get('sm_max_mhz') or 1965.0)
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')      # dram bytes/launch from the last ncu --set full capture
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = (json.load(f).get(dom) or {}).get('dram_bytes_per_launch')
        roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                    'frac': achieved / peaks['hbm_gbs'], 'peak_source': f'{peak_kind} copy bandwidth (MEASURED_PEAKS.json)',
                    'traffic': traffic, 'algorithmic_bytes_per_launch': alg_bytes, 'launch_ms': dom_ms,
                    'note': 'per-edge k/v tensors are never materialised; the kernel streams node planes + the R-cache; bytes are '
                            'counted for the rows a launch really processes (receptive-field pruning); the loaded unit is the '
                            'L1/shared data pipe, not DRAM (DESIGN.md section 5)',
                    'rows_per_launch': rows,
                    'fp32': {'achieved_tflops': flops_node * rows / (dom_ms * 1e-3) / 1e12, 'peak_tflops': fp32_peak,
                             'frac': flops_node * rows / (dom_ms * 1e-3) / 1e12 / fp32_peak,
                             'peak_source': 'nominal 148 SM x 128 FMA x 2 x max SM clock'}}

    # ---- end to end through the public API: host batch -> model.sample() -> host trajectory -------
    e2e = None
    if not args.no_e2e:
        host_batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
        h2d = sum(v.numel() * v.element_size() for v in host_batch.values() if torch.is_tensor(v))
        steps_e2e = min(args.e2e_steps, T_STEPS)
        # warm-up of the public-API path (allocations, R-cache buffers, NCCL channels of the final gather)
        model.sample(host_batch, num_steps=2, traj_mode='final')
        if world > 1:
            from cbgbench_b200 import sharding
            gid0 = (batch['ligand_element_batch'] + rank * B).to(dev)
            sharding.gather_final(torch.zeros(n_lig_tot, 3, device=dev), torch.zeros(n_lig_tot, dtype=torch.int64, device=dev),
                                  gid0, counts=[n_lig_tot] * world)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        traj = model.sample(host_batch, num_steps=steps_e2e)          # H2D + steps + D2H of the trajectory
        t_last = T_STEPS - steps_e2e
        x_fin, v_fin = traj[t_last][0], traj[t_last][1].argmax(-1)    # what sample.py consumes (traj[0] at full T)
        if world > 1:                                                  # the single gather of final coordinates
            from cbgbench_b200 import sharding
            gid = (batch['ligand_element_batch'] + rank * B).to(dev)
            sharding.gather_final(x_fin.to(dev), v_fin.to(dev), gid, counts=[n_lig_tot] * world)
        ev1.record()
        barrier()
        e2e_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        d2h = sum(traj[t][0].numel() * 4 + traj[t][1].numel() * 4 for t in traj if t >= 0)
        e2e_value = (B * world) / (float(e2e_ms.item()) * 1e-3) * (steps_e2e / T_STEPS)
        e2e = {'value': e2e_value, 'unit': 'ligands/s', 'h2d_bytes_per_step': h2d / steps_e2e,
               'd2h_bytes_per_step': d2h / steps_e2e, 'seconds': float(e2e_ms.item()) * 1e-3, 'denoise_steps': steps_e2e,
               'api': 'TargetDiffB200.sample(host batch) -> traj (CPU), H2D/D2H and final gather inside the timed region'}

    # ---- CPU baseline beside it (rank 0, N = 1) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_graphs, threads, tried = pick_cpu_sample(args.workload, budget_s=6.0)
        sec = cpu_reference_steps(args.workload, n_graphs, 2, 1, threads)
        cpu = {'value': n_graphs / (T_STEPS * sec), 'unit': 'ligands/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n_graphs} of {B} pockets ({n_prot}+{n_lig} atoms), 2 denoise steps after 1 warm-up, '
                         f'oracle port (torch CPU fp32), extrapolated: pockets / (1000 x s/step); {threads} torch threads '
                         f'of {os.cpu_count()} host cores (fastest of {tried} s/step for 1 pocket)'}

    if rank == 0:
        line = {
            'metric': 'ligands/sec sampled (1000-step denoise, batch 64)', 'value': value, 'unit': 'ligands/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{args.workload}: {desc}', 'graphs_per_gpu': B, 'nodes_per_gpu': N,
                       'denoise_steps_per_ligand': T_STEPS, 'step': 'one denoise step of the whole batch',
                       'l2': 'flushed (256 MiB write) between timed steps', 'parallelism': f'dp{world} (pockets sharded, no data-path collective)'},
            'clocks': clocks.summary(), 'gpu_launches': int(gpu_launches), 'e2e': e2e, 'roofline': roofline,
            'cpu_baseline': cpu, 'kernels': kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
