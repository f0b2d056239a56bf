"""ORACLE (test infrastructure only, never imported by the product): CPU restatement of the reference's
sampling-time transforms and batch construction.

What ``sample.py:177-183`` does per pocket: ``dataset[i]`` is evaluated ``num_samples`` times, each time running
the transform list of ``configs/*/test/*.yml`` on the raw pocket, then PyG's ``DataLoader`` collates the samples.
This module restates those transforms with every random draw INJECTED (the reference draws from numpy's and
torch's global generators), so that the same numbers can be fed to the CUDA batch builder:

* ``space_size``                init_lig.py:247-250   (pdist, sort, median of the 10 largest distances)
* ``choose_num_atoms``          init_lig.py:27-31,47-52 + numpy's legacy ``RandomState.choice(a, p=p)``
                                 (cdf = cumsum(p); cdf /= cdf[-1]; idx = searchsorted(cdf, u, 'right'))
* ``featurize_protein``         protein_featurizer.py:19-30
* ``uniform_types``             init_lig.py:22-26,385-388 (Gumbel arg-max over zero logits)
* ``denovo_sample``             configs/denovo/test/{targetdiff,diffbp,diffsbdd}.yml:
                                 center_pos / center_whole_pos (translation.py:5-50) -> assign_molsize
                                 (init_lig.py:232-250) -> assign_atomtype (:373-401) -> assign_molpos (:404-421)
* ``context_sample``            configs/{linker,frag,scaffold,sidechain}/test/*.yml: assign_gensize
                                 (init_lig.py:253-296) -> assign_genatomtype (:299-341) -> center_pos with
                                 mask_flag=ctx_flag (translation.py:11-24) -> assign_genpos (:441-457)
* ``collate``                   merge.py:6-25 (key prefixing) + PyG ``Batch.from_data_list`` with ``follow_batch``
                                 (third-party, not vendored: concatenation along dim 0 and ``<key>_batch`` vectors;
                                 this part of the restatement is a definition, "parity unpinned")

Pinned against the live reference by ``tests/golden/make_golden_f3.py`` (which imports the unmodified transform
classes, replays their random draws and asserts equality with this module).
"""
import numpy as np
import torch
import torch.nn.functional as F

ATOMIC_NUMBERS = (1, 6, 7, 8, 16, 34)     # repo/utils/protein/constants.py (H, C, N, O, S, Se)


def space_size(pos):
    """init_lig.py:247-250."""
    d = torch.pdist(pos)
    d = torch.sort(d, descending=True)[0]
    return torch.median(d[:10])


def bin_index(size, bounds):
    """init_lig.py:47-52 (python floats: the reference passes ``pocket_size.item()``)."""
    for i, b in enumerate(bounds):
        if b > size:
            return i
    return len(bounds)


def choose_num_atoms(size, table, u):
    """sample_atom_num (init_lig.py:27-31) with the uniform draw of ``np.random.choice`` injected."""
    values, probs = table['bins'][bin_index(size, table['bounds'])]
    cdf = np.cumsum(np.asarray(probs, dtype=np.float64))
    cdf /= cdf[-1]
    return int(np.asarray(values)[int(np.searchsorted(cdf, u, side='right'))])


def featurize_protein(element, atom_to_aa_type, is_backbone):
    """FeaturizeProteinFullAtom (protein_featurizer.py:19-30): [6 element one-hot | backbone flag], aa type."""
    an = torch.tensor(ATOMIC_NUMBERS, dtype=torch.long)
    onehot = (element.view(-1, 1) == an.view(1, -1)).float()
    feat = torch.cat([onehot, is_backbone.view(-1, 1).long()], dim=-1)      # float + long -> float (torch.cat promotion)
    return feat, atom_to_aa_type


def uniform_types(u):
    """log_sample_categorical over zero logits (init_lig.py:22-26) with the uniform tensor [n,K] injected."""
    gumbel = -torch.log(-torch.log(u + 1e-30) + 1e-30)
    return (gumbel + torch.zeros_like(u)).argmax(dim=-1)


def denovo_sample(prot_pos, table, u_size, type_dist, type_u, pos_dist, pos_noise, num_classes):
    """One de-novo sample.  Returns (protein_pos_centred, centre[1,3], ligand_pos, ligand_atom_type)."""
    centre = prot_pos.mean(dim=0, keepdim=True)                       # center_pos(protein) / center_whole_pos without ligand
    p = prot_pos - centre
    n = choose_num_atoms(space_size(p).item(), table, u_size)         # assign_molsize runs on the centred pocket
    if type_dist == 'uniform':
        t = uniform_types(type_u[:n]).long()
    elif type_dist == 'absorbing':
        t = torch.zeros(n, dtype=torch.long)                          # absorbing_state = 0 (constants.py)
    elif type_dist == 'zeros':
        t = torch.zeros(n, num_classes, dtype=torch.long)
    else:
        raise ValueError(type_dist)
    x = pos_noise[:n].clone()
    if pos_dist == 'zero_mean_gaussian':
        x -= torch.mean(x, dim=0, keepdim=True)
    elif pos_dist != 'gaussian':
        raise ValueError(pos_dist)
    return p, centre, x, t


def context_sample(prot_pos, ctx_pos, ctx_type, table, u_size, extra, type_u, pos_noise):
    """One linker / fragment / scaffold / side-chain sample: fixed context atoms first, generated atoms after.
    extra = the ``torch.randint(1, 8)`` draw (used only when the prior asks for no more atoms than the context has).
    Returns (protein_pos_centred, centre, ligand_pos, ligand_atom_type, ctx_flag)."""
    n = choose_num_atoms(space_size(prot_pos).item(), table, u_size)
    c = ctx_pos.shape[0]
    used_extra = n <= c
    if used_extra:
        n = c + int(extra)
    pos = torch.zeros(n, 3)
    pos[:c] = ctx_pos
    t = torch.zeros(n, dtype=torch.long)
    t[:c] = ctx_type
    ctx = torch.zeros(n, dtype=torch.bool)
    ctx[:c] = True
    gen = ~ctx
    t = torch.where(gen, uniform_types(type_u[:n]), t).long()         # assign_genatomtype
    centre = pos[ctx].mean(dim=0, keepdim=True) if ctx.sum() > 0 else pos.mean(dim=0, keepdim=True)
    p = prot_pos - centre
    pos = pos - centre
    pos = torch.where(gen.unsqueeze(-1), pos_noise[:n], pos)          # assign_genpos (gaussian)
    return p, centre, pos, t, ctx, used_extra


def collate(samples, feat, aa):
    """samples: list of dicts with protein_pos, centre, ligand_pos, ligand_atom_type[, ctx_flag]; the protein features
    are the same for every sample of a pocket.  Returns the flat batch the samplers consume."""
    P = feat.shape[0]
    out = {
        'protein_pos': torch.cat([s['protein_pos'] for s in samples]),
        'protein_atom_feature': feat.repeat(len(samples), 1),
        'protein_aa_type': aa.repeat(len(samples)),
        'protein_lig_flag': torch.zeros(P * len(samples), dtype=torch.bool),
        'protein_element_batch': torch.arange(len(samples)).repeat_interleave(P),
        'protein_translation': torch.cat([s['centre'].expand(P, -1) for s in samples]),
        'ligand_pos': torch.cat([s['ligand_pos'] for s in samples]),
        'ligand_atom_type': torch.cat([s['ligand_atom_type'] for s in samples]),
        'ligand_element_batch': torch.cat([torch.full((s['ligand_pos'].shape[0],), i, dtype=torch.long)
                                           for i, s in enumerate(samples)]),
    }
    out['ligand_lig_flag'] = torch.ones(out['ligand_pos'].shape[0], dtype=torch.bool)
    if 'ctx_flag' in samples[0]:
        out['ligand_ctx_flag'] = torch.cat([s['ctx_flag'] for s in samples])
        out['ligand_gen_flag'] = ~out['ligand_ctx_flag']
    return out
