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
from cbgbench_b200.diffbp import DiffBPB200
from cbgbench_b200.schedulers import DiffsbddVariationalTables
from oracle import diffusion_sbdd as OS, diffusion_bp as OB
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


# =============================================================================================================
# DiffBP
# =============================================================================================================

def bp_model(num_steps=10, device=None, **kw):
    model = DiffBPB200(synthetic.diffbp_config(num_steps=num_steps, **kw))
    sd = synthetic.seeded_state_dict(model, seed=WEIGHT_SEED)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return (model.to(device) if device is not None else model), sd


def bp_batch(n_prot, n_lig, seed, gen_mode='denovo'):
    """Ligand types start at the absorbing state 0, every 5th atom keeps a random type (= make_golden_f2.bp_batch)."""
    batch = synthetic.make_batch(n_prot, n_lig, seed=seed, gen_mode=gen_mode)
    v = batch['ligand_atom_type'].clone()
    keep = torch.arange(v.numel()) % 5 == 4
    batch['ligand_atom_type'] = torch.where(keep, v, torch.zeros_like(v))
    return batch


def test_bp_state_dict_keys_match_reference():
    with open(os.path.join(GOLDEN, 'bp_state_keys.json')) as f:
        want = json.load(f)
    model, _ = bp_model(10)
    have = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert list(have.keys()) == list(want.keys())
    assert have == want


def test_bp_oracle_trajectory_matches_reference():
    g = golden('bp_trajectory.npz')
    T = 10
    _, sd = bp_model(T)
    batch = bp_batch([150, 60], [20, 9], seed=41)
    pn, tu = synthetic.make_bp_noise(T, 29, seed=13)
    traj = OB.sample(sd, batch, T, pn, tu)
    for t in range(-1, T):
        assert np.array_equal(traj[t][1].argmax(-1).numpy(), g[f'v{t}']), t
        assert rel_err(traj[t][0], g[f'x{t}']) < 1e-5, t
    eps, eps_com, _ = OB.denoise(sd, batch, batch['ligand_pos'].float(),
                                 torch.nn.functional.one_hot(batch['ligand_atom_type'], 13).float())
    assert rel_err(eps + eps_com, g[f'eps{T - 1}']) < 1e-5


def test_bp_mask_type_step_properties():
    """MaskTypeSchedule.backward_remove_noise: only generated absorbing-state atoms may change; prob = (T - t) / T."""
    rs = np.random.RandomState(0)
    n, K, T = 50, 13, 20
    logits = torch.from_numpy(rs.normal(size=(n, K)).astype(np.float32))
    v = torch.from_numpy(rs.randint(0, 3, size=n))
    c = torch.nn.functional.one_hot(v, K).float()
    gen = torch.from_numpy(rs.rand(n) < 0.7)
    u = torch.from_numpy(rs.random_sample(n).astype(np.float32))
    for t in (0, 7, 19):
        cn, vn = OB.mask_type_reverse_step(logits, c, t, T, gen, u, K)
        changed = vn != v
        assert not bool((changed & ~(gen & (v == 0))).any())
        want = gen & (v == 0) & (u < (T - t) / T) & (logits.argmax(-1) != 0)
        assert torch.equal(changed, want)
    model, _ = bp_model(T)
    assert model.type_scheduler.change_prob(0) == 1.0 and abs(model.type_scheduler.change_prob(19) - 0.05) < 1e-7


@pytest.mark.gpu
def test_bp_trajectory_matches_golden_and_oracle():
    g = golden('bp_trajectory.npz')
    T = 10
    model, sd = bp_model(T, device='cuda')
    batch = bp_batch([150, 60], [20, 9], seed=41)
    pn, tu = synthetic.make_bp_noise(T, 29, seed=13)
    eps = {}
    traj = model.sample(batch, pos_noise=pn, type_uniform=tu, eps_out=eps)
    assert sorted(traj.keys()) == list(range(-1, T))
    assert traj[-1][0].is_cuda and not traj[0][0].is_cuda
    for t in range(-1, T):
        assert np.array_equal(traj[t][1].argmax(-1).cpu().numpy(), g[f'v{t}']), t
        assert rel_err(traj[t][0].cpu(), g[f'x{t}']) < TOL, t
    for t in range(T):
        assert rel_err(eps[t].cpu(), g[f'eps{t}']) < TOL, t


@pytest.mark.gpu
@pytest.mark.parametrize('case', [
    dict(n_prot=[40, 33, 20], n_lig=[9, 6, 4], T=6, layers=3, com=3, gen_mode='denovo'),
    dict(n_prot=[60, 10], n_lig=[12, 30], T=5, layers=2, com=2, gen_mode='partial'),
    dict(n_prot=[5, 0, 70], n_lig=[3, 6, 10], T=4, layers=2, com=1, gen_mode='denovo'),
], ids=['ragged', 'partial_gen', 'no_pocket'])
def test_bp_sample_matches_oracle(case):
    T = case['T']
    model, sd = bp_model(T, device='cuda', num_layers=case['layers'], num_layers_com=case['com'])
    batch = bp_batch(case['n_prot'], case['n_lig'], seed=19, gen_mode=case['gen_mode'])
    n_lig = int(sum(case['n_lig']))
    pn, tu = synthetic.make_bp_noise(T, n_lig, seed=5)
    want = OB.sample(sd, batch, T, pn, tu)
    for rcache, prune in ((True, True), (False, False)):
        model.use_rcache, model.use_prune = rcache, prune
        traj = model.sample(batch, pos_noise=pn, type_uniform=tu)
        for t in range(-1, T):
            assert torch.equal(traj[t][1].argmax(-1).cpu(), want[t][1].argmax(-1)), (t, rcache)
            assert rel_err(traj[t][0].cpu(), want[t][0]) < TOL, (t, rcache)


@pytest.mark.gpu
def test_bp_com_shift_is_measurable():
    """The CoM head must matter in the comparison above: eps + eps_com differs from eps alone by more than TOL."""
    T = 4
    model, sd = bp_model(T, device='cuda', num_layers=2)
    batch = bp_batch([40, 30], [9, 7], seed=23)
    x = batch['ligand_pos'].float()
    c = torch.nn.functional.one_hot(batch['ligand_atom_type'], 13).float()
    eps, eps_com, _ = OB.denoise(sd, batch, x, c)
    assert float(eps_com.abs().max()) > 10 * TOL * float(eps.abs().max())
    got = {}
    pn, tu = synthetic.make_bp_noise(T, 16, seed=1)
    model.sample(batch, pos_noise=pn, type_uniform=tu, num_steps=1, eps_out=got)
    assert rel_err(got[T - 1].cpu(), eps + eps_com) < TOL
    assert rel_err(got[T - 1].cpu(), eps) > 10 * TOL


@pytest.mark.gpu
def test_bp_free_running_sample_reproducible():
    T = 6
    model, _ = bp_model(T, device='cuda', num_layers=2)
    batch = bp_batch([50, 40], [10, 8], seed=4)
    torch.manual_seed(3)
    a = model.sample(batch, traj_mode='final')
    torch.manual_seed(3)
    b = model.sample(batch, traj_mode='final')
    assert torch.isfinite(a[0][0]).all()
    assert torch.equal(a[0][0], b[0][0]) and torch.equal(a[0][1], b[0][1])
    assert torch.equal(a[-1][0], b[-1][0])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['diffsbdd', 'diffbp'])
def test_sample_driver_runs_f2_models(name):
    """The sample.py-style driver (row f1) with the row-f2 models: per-pocket results, reproducible from the seed."""
    from cbgbench_b200 import sample_driver
    argv = ['--model', name, '--pockets', '3', '--batch-size', '2', '--n-prot', '30', '--n-lig', '5', '--steps', '4',
            '--layers', '2']
    a = sample_driver.main(argv)
    b = sample_driver.main(argv)
    assert len(a) == 3
    for ra, rb in zip(a, b):
        assert ra['pos'].shape == (5, 3) and ra['v'].shape == (5,)
        assert torch.isfinite(ra['pos']).all() and int(ra['v'].max()) < 13
        assert torch.equal(ra['pos'], rb['pos']) and torch.equal(ra['v'], rb['v'])
