"""the DiffSBDD / DiffBP samplers on the same denoiser kernels.

CPU part: the oracle restatements against the reference's own outputs (tests/golden/*_trajectory.npz, generated
by tests/golden/make_golden_f2.py from the unmodified reference) and the host-side mirrors (state-dict keys,
schedule tables).  GPU part (-m gpu): the CUDA path against the oracle and the golden fixtures."""
import json
import os

import numpy as np
import pytest
import torch

from cbgbench_b200 import synthetic
from cbgbench_b200.diffsbdd import DiffSBDDB200
from cbgbench_b200.schedulers import DiffsbddVariationalTables
from oracle import diffusion_sbdd as OS
from helpers import GOLDEN, WEIGHT_SEED, golden, rel_err

torch.set_grad_enabled(False)
TOL = 1e-4            # north-star: 1e-4 relative fp32


def sbdd_model(num_steps=10, device=None, **kw):
    model = DiffSBDDB200(synthetic.diffsbdd_config(num_steps=num_steps, **kw))
    sd = synthetic.seeded_state_dict(model, seed=WEIGHT_SEED)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return (model.to(device) if device is not None else model), sd


# ---- CPU: oracle + host logic -------------------------------------------------------------------------------

def test_sbdd_state_dict_keys_match_reference():
    with open(os.path.join(GOLDEN, 'sbdd_state_keys.json')) as f:
        want = json.load(f)
    model, _ = sbdd_model(10)
    have = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert list(have.keys()) == list(want.keys())
    assert have == want


def test_sbdd_gamma_tables_match_reference():
    g = golden('sbdd_trajectory.npz')
    for T in (10, 1000):
        tab = DiffsbddVariationalTables(T, 'polynomial_2')
        assert np.array_equal(tab.gamma.gamma.numpy(), g[f'gamma_T{T}'])
        assert np.array_equal(OS.gamma_table(T).numpy(), g[f'gamma_T{T}'])
    with pytest.raises(NotImplementedError):
        DiffsbddVariationalTables(10, 'cosine')


def test_sbdd_step_scalars_match_oracle():
    tab = DiffsbddVariationalTables(1000, 'polynomial_2')
    gamma = tab.gamma.gamma.detach()
    for t in (0, 1, 500, 999):
        want = [float(v) for v in OS.step_scalars(gamma, t, 1000)]
        assert list(tab.step_scalars(t)) == want
    a0, s0, sx = OS.final_scalars(gamma, 1000)
    inv_a, b, s = tab.final_scalars()
    assert inv_a == float(1.0 / a0) and b == float(s0) and s == float(sx)


def test_sbdd_oracle_trajectory_matches_reference():
    g = golden('sbdd_trajectory.npz')
    T = 10
    _, sd = sbdd_model(T)
    batch = synthetic.make_batch([150, 60], [20, 9], seed=31)
    noise = synthetic.make_sbdd_noise(T, 29, 13, seed=9)
    traj, _ = OS.sample(sd, batch, T, noise)
    for t in range(-1, T):
        assert rel_err(traj[t][0], g[f'x{t}']) < 1e-5, t
        assert rel_err(traj[t][1], g[f'c{t}']) < 1e-5, t


def test_sbdd_oracle_keeps_ligand_com_at_zero():
    """Size-independent property of the path: every state after a COM projection has zero ligand mean per graph."""
    T = 4
    _, sd = sbdd_model(T, num_layers=2)
    batch = synthetic.make_batch([30, 25, 12], [7, 5, 3], seed=5)
    noise = synthetic.make_sbdd_noise(T, 15, 13, seed=2)
    traj, _ = OS.sample(sd, batch, T, noise)
    bl = batch['ligand_element_batch']
    for t in traj:
        for gidx in range(3):
            assert float(traj[t][0][bl == gidx].mean(0).abs().max()) < 1e-5


# ---- GPU: CUDA path vs oracle / golden ----------------------------------------------------------------------

@pytest.mark.gpu
def test_sbdd_trajectory_matches_golden_and_oracle():
    g = golden('sbdd_trajectory.npz')
    T = 10
    model, sd = sbdd_model(T, device='cuda')
    batch = synthetic.make_batch([150, 60], [20, 9], seed=31)
    noise = synthetic.make_sbdd_noise(T, 29, 13, seed=9)
    traj = model.sample(batch, noise=noise)
    assert sorted(traj.keys()) == list(range(-1, T))
    assert traj[-1][0].is_cuda and not traj[0][0].is_cuda and not traj[T - 1][0].is_cuda
    for t in range(-1, T):
        assert rel_err(traj[t][0].cpu(), g[f'x{t}']) < TOL, t
        assert rel_err(traj[t][1].cpu(), g[f'c{t}']) < TOL, t
    assert model.last_launches > 0


@pytest.mark.gpu
@pytest.mark.parametrize('case', [
    dict(n_prot=[40, 33, 20], n_lig=[9, 6, 4], T=6, layers=3, gen_mode='denovo'),
    dict(n_prot=[60, 10], n_lig=[12, 30], T=5, layers=2, gen_mode='partial'),
    dict(n_prot=[5, 0, 70], n_lig=[3, 6, 10], T=4, layers=2, gen_mode='denovo'),      # a graph without pocket atoms
], ids=['ragged', 'partial_gen', 'no_pocket'])
def test_sbdd_sample_matches_oracle(case):
    T = case['T']
    model, sd = sbdd_model(T, device='cuda', num_layers=case['layers'])
    batch = synthetic.make_batch(case['n_prot'], case['n_lig'], seed=17, gen_mode=case['gen_mode'])
    n_lig = int(sum(case['n_lig']))
    noise = synthetic.make_sbdd_noise(T, n_lig, 13, seed=3)
    want, _ = OS.sample(sd, batch, T, noise)
    traj = model.sample(batch, noise=noise)
    for t in range(-1, T):
        assert rel_err(traj[t][0].cpu(), want[t][0]) < TOL, t
        assert rel_err(traj[t][1].cpu(), want[t][1]) < TOL, t
    # early stop: no final stage, traj[t_last - 1] on the device
    part = model.sample(batch, noise=noise, num_steps=2, traj_mode='final')
    assert sorted(part.keys()) == [T - 3, T - 2]
    assert rel_err(part[T - 3][0].cpu(), want[T - 3][0]) < TOL


@pytest.mark.gpu
def test_sbdd_free_running_sample_is_finite_and_centred():
    """torch-drawn noise (the production path): finite output, zero ligand COM per graph, seed-reproducible."""
    T = 8
    model, _ = sbdd_model(T, device='cuda', num_layers=2)
    batch = synthetic.make_batch([50, 40], [10, 8], seed=4)
    torch.manual_seed(11)
    a = model.sample(batch, traj_mode='final')
    torch.manual_seed(11)
    b = model.sample(batch, traj_mode='final')
    x, c, bl = a[0]
    assert torch.isfinite(x).all() and torch.isfinite(c).all()
    assert torch.equal(x, b[0][0]) and torch.equal(c, b[0][1])
    for gidx in range(2):
        assert float(x[bl == gidx].mean(0).abs().max()) < 1e-4


@pytest.mark.gpu
def test_sbdd_rejects_rcache_plan():
    """The C-ABI refuses a plan with an R-cache (the pocket moves)."""
    import ctypes as C
    from cbgbench_b200 import _lib
    from helpers import make_model
    model, _ = make_model(4, device='cuda', num_layers=1)
    batch = synthetic.make_batch([20], [5], seed=1)
    model.use_rcache = True
    state = model.prepare(batch)
    if not state['plan'].rcache:
        pytest.skip('R-cache disabled in this environment')
    coef = _lib.SbddCoef(a=1.0, b=0.0, s=0.0, mode=0)
    z = torch.zeros(5 * 13, device='cuda')
    rc = _lib.lib().cbg_sbdd_step_f32(C.byref(state['plan']), C.byref(coef), z.data_ptr(), z.data_ptr(), z.data_ptr(),
                                      z.data_ptr(), z.data_ptr(), z.data_ptr(), None, None, None)
    assert rc != 0 and b'R-cache' in _lib.lib().cbg_last_error()
