"""Golden vectors for SURVEY.md section 8 row f3 (sampling-time transforms + batch construction).

Runs the UNMODIFIED reference transform classes (repo/datasets/transforms/{protein_featurizer,translation,init_lig}.py)
on synthetic raw pockets in this container, replays the random draws they consumed (numpy legacy generator for the
size prior, torch's CPU generator for types / positions), checks that ``oracle/batch_builder.py`` reproduces every
sample, and stores inputs, draws and the reference outputs (plus the reference's size-prior table
``_atom_num_dist.npy``, which is data the transforms load) in ``tests/golden/batch_builder.npz``.

    python tests/golden/make_golden_f3.py        (needs /root/reference; the fixture is committed)
"""
import importlib
import os
import sys
import types
import typing

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402
from oracle import batch_builder as OB  # noqa: E402


def load_reference():
    ref_shims.install()
    torch.Any = typing.Any            # init_lig.py annotates with torch.Any, which torch 2.11 no longer has
    for pkg in ['repo.utils', 'repo.utils.molecule', 'repo.utils.protein', 'repo.models.utils']:
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(ref_shims.REF_ROOT, *pkg.split('.'))]
            sys.modules[pkg] = m
    il = importlib.import_module('repo.datasets.transforms.init_lig')
    tr = importlib.import_module('repo.datasets.transforms.translation')
    pf = importlib.import_module('repo.datasets.transforms.protein_featurizer')
    return il, tr, pf


def raw_pocket(rs, n_atoms, sigma, shift):
    el = rs.choice([1, 6, 7, 8, 16], size=n_atoms, p=[0.05, 0.55, 0.18, 0.2, 0.02])   # no Se: the reference's protein atom_type map has no 34
    return {
        'element': torch.from_numpy(el.astype(np.int64)),
        'pos': torch.from_numpy((rs.normal(0, sigma, size=(n_atoms, 3)) + shift).astype(np.float32)),
        'atom_to_aa_type': torch.from_numpy(rs.randint(0, 20, size=n_atoms).astype(np.int64)),
        'is_backbone': torch.from_numpy(rs.randint(0, 2, size=n_atoms).astype(bool)),
        'atom_name': ['CA' if i % 7 == 0 else 'C' for i in range(n_atoms)],
    }


def main():
    il, tr, pf = load_reference()
    ED = ref_shims.EasyDict
    TD = importlib.import_module('repo.datasets.transforms._base').TRANSFORM_DICT
    table = il.config_atom_num
    out = {'table/bounds': np.asarray(table['bounds'], dtype=np.float64),
           'table/bin_ptr': np.cumsum([0] + [len(b[0]) for b in table['bins']]).astype(np.int32),
           'table/values': np.concatenate([np.asarray(b[0], dtype=np.int32) for b in table['bins']]),
           'table/probs': np.concatenate([np.asarray(b[1], dtype=np.float64) for b in table['bins']])}
    tab = {'bounds': list(table['bounds']), 'bins': [(list(b[0]), list(b[1])) for b in table['bins']]}
    rs = np.random.RandomState(7)
    n_samples = 6
    # (name, recipe, protein atoms, sigma [A], type distribution, mode, position distribution, context atoms)
    cases = [('denovo_targetdiff', 'denovo', 310, 4.6, 'uniform', 'add_aromatic', 'gaussian', 0),
             ('denovo_diffbp', 'denovo', 180, 5.4, 'absorbing', 'add_aromatic', 'gaussian', 0),
             ('denovo_diffsbdd', 'denovo', 420, 10.4, 'zeros', 'basic', 'zero_mean_gaussian', 0),
             ('linker', 'context', 260, 5.0, 'uniform', 'add_aromatic', 'gaussian', 14),
             ('scaffold_big_ctx', 'context', 350, 5.3, 'uniform', 'add_aromatic', 'gaussian', 58),
             ('frag_no_ctx', 'context', 90, 5.2, 'uniform', 'add_aromatic', 'gaussian', 0)]
    names = []
    for ci, (name, recipe, n_prot, sigma, type_dist, mode, pos_dist, n_ctx) in enumerate(cases):
        raw = raw_pocket(rs, n_prot, sigma, rs.normal(0, 20, size=3))
        ctx_pos = torch.from_numpy((rs.normal(0, 2.0, size=(n_ctx, 3)) + raw['pos'].mean(0).numpy()).astype(np.float32))
        ctx_el = torch.from_numpy(rs.choice([6, 7, 8], size=n_ctx).astype(np.int64))
        K = len(il.map_atom_type_aromatic_to_index) if mode == 'add_aromatic' else len(il.map_atom_type_only_to_index)
        ctx_type = torch.from_numpy(rs.randint(0, K, size=n_ctx).astype(np.int64))
        seed = 100 + ci
        np.random.seed(seed)
        torch.manual_seed(seed)
        ref_samples = []
        for s in range(n_samples):                                   # sample.py:177: dataset[i] evaluated num_samples times
            data = ED({'protein': dict(raw)})
            data = pf.FeaturizeProteinFullAtom()(data)
            if recipe == 'denovo':
                data.ligand = {}                                      # remove_ligand (molecule_featurizer.py:163-171)
                data = TD['center_pos'](center_flag='protein')(data) if type_dist != 'zeros' else \
                    TD['center_whole_pos']()(data)
                data = il.AssignMolSize('prior_distcond')(data)
                data = il.AssignMolType(type_dist, mode)(data)
                data = il.AssignMolPos(pos_dist)(data)
            else:
                # state after choose_ctx_gen + remove_ligand_gen (molecule_featurizer.py:173-195): context atoms only
                data.ligand = {'atom_type': ctx_type.clone(), 'element': ctx_el.clone(), 'pos': ctx_pos.clone(),
                               'ctx_flag': torch.ones(n_ctx, dtype=torch.bool), 'gen_flag': torch.zeros(n_ctx, dtype=torch.bool),
                               'lig_flag': torch.ones(n_ctx, dtype=torch.bool)}
                data = il.AssignGenSize('prior_distcond')(data)
                data = il.AssignGenType(type_dist, mode)(data)
                data = TD['center_pos'](center_flag='ligand', mask_flag='ctx_flag')(data)
                data = il.AssignGenPos(pos_dist)(data)
            ref_samples.append(data)
        # ---- replay the draws
        np.random.seed(seed)
        torch.manual_seed(seed)
        feat, aa = OB.featurize_protein(raw['element'], raw['atom_to_aa_type'], raw['is_backbone'])
        u_size, extras, type_us, noises, ora = [], [], [], [], []
        for s in range(n_samples):
            ref = ref_samples[s]
            n = int(ref.ligand.pos.shape[0])
            u = np.random.random_sample()
            u_size.append(u)
            tu = torch.zeros(n, K)
            if recipe == 'denovo':
                if type_dist == 'uniform':
                    tu = torch.rand(n, K)
                pn = torch.randn(n, 3)
                p, centre, x, t = OB.denovo_sample(raw['pos'], tab, u, type_dist, tu, pos_dist, pn, K)
                extras.append(0)
                ctx = None
            else:
                # torch.randint is consumed only when the prior asks for <= ctx atoms; probe with a generator copy
                state = torch.get_rng_state()
                n_prior = OB.choose_num_atoms(OB.space_size(raw['pos']).item(), tab, u)
                ex = int(torch.randint(1, 8, size=(1,))) if n_prior <= n_ctx else 0
                if n_prior > n_ctx:
                    torch.set_rng_state(state)
                extras.append(ex)
                tu = torch.rand(n, K)
                pn = torch.randn(n, 3)
                p, centre, x, t, ctx, _ = OB.context_sample(raw['pos'], ctx_pos, ctx_type, tab, u, ex, tu, pn)
            type_us.append(tu)
            noises.append(pn)
            # oracle == reference, sample by sample
            assert x.shape[0] == n, (name, s, x.shape, n)
            assert torch.equal(p, ref.protein.pos), (name, s)
            assert torch.equal(centre.expand(n_prot, -1), ref.protein.translation), (name, s)
            assert torch.equal(x, ref.ligand.pos), (name, s)
            assert torch.equal(t, ref.ligand.atom_type), (name, s, t, ref.ligand.atom_type)
            assert torch.equal(feat, ref.protein.atom_feature) and torch.equal(aa, ref.protein.aa_type)
            if ctx is not None:
                assert torch.equal(ctx, ref.ligand.ctx_flag) and torch.equal(~ctx, ref.ligand.gen_flag)
            smp = {'protein_pos': p, 'centre': centre, 'ligand_pos': x, 'ligand_atom_type': t}
            if ctx is not None:
                smp['ctx_flag'] = ctx
            ora.append(smp)
        batch = OB.collate(ora, feat, aa)
        names.append(name)
        pre = f'{name}/'
        out[pre + 'meta'] = np.array([n_prot, n_ctx, n_samples, K, {'uniform': 0, 'absorbing': 1, 'zeros': 2}[type_dist],
                                      {'gaussian': 0, 'zero_mean_gaussian': 1}[pos_dist], 0 if recipe == 'denovo' else 1])
        out[pre + 'raw_pos'] = raw['pos'].numpy()
        out[pre + 'raw_element'] = raw['element'].numpy()
        out[pre + 'raw_aa'] = raw['atom_to_aa_type'].numpy()
        out[pre + 'raw_backbone'] = raw['is_backbone'].numpy()
        out[pre + 'ctx_pos'] = ctx_pos.numpy()
        out[pre + 'ctx_type'] = ctx_type.numpy()
        out[pre + 'u_size'] = np.asarray(u_size, dtype=np.float64)
        out[pre + 'extra'] = np.asarray(extras, dtype=np.int32)
        out[pre + 'type_u'] = torch.cat(type_us).numpy()
        out[pre + 'pos_noise'] = torch.cat(noises).numpy()
        out[pre + 'space_size'] = np.float32(OB.space_size(raw['pos']).item())
        for k, v in batch.items():
            out[pre + 'batch/' + k] = v.numpy()
        sizes = [int(s['ligand_pos'].shape[0]) for s in ora]
        print(f'{name}: sizes {sizes} extras {extras} space {float(out[pre + "space_size"]):.4f} '
              f'bin {OB.bin_index(float(out[pre + "space_size"]), tab["bounds"])}')
    out['names'] = np.array(names)
    path = os.path.join(HERE, 'batch_builder.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
