"""Device-side construction of sampling batches.

The reference's ``sample.py:177`` evaluates ``dataset[i]`` ``num_samples`` times per pocket - 100-200 passes of the Python
transform list of ``configs/<task>/test/<model>.yml`` - and collates the samples with PyG.  ``DeviceBatchBuilder`` takes
the same transform list (``from_transform_cfg``), keeps the raw pocket on the GPU and produces the collated batch with
three kernels (``csrc/batch.cu``): pocket centre + space size, size prior, per-sample fill.

Transforms understood (names and options of repo/datasets/transforms):
    featurize_protein_fa                                       protein_featurizer.py:7-42
    remove_ligand | choose_ctx_gen + remove_ligand_gen         (the caller passes the context atoms, if any)
    center_pos(center_flag=protein) | center_whole_pos | center_pos(center_flag=ligand, mask_flag=ctx_flag)
    assign_molsize | assign_gensize (prior_distcond)           init_lig.py:232-296
    assign_atomtype | assign_genatomtype (uniform | absorbing | zeros)   init_lig.py:299-401
    assign_molpos | assign_genpos (gaussian | zero_mean_gaussian)        init_lig.py:404-457
    merge                                                      merge.py:6-25 (key prefixes ``protein_`` / ``ligand_``)
Anything else raises (no silent approximation, no CPU fallback).

Random numbers are drawn with torch on the device in the reference's per-sample order of KINDS (size uniform, optional
randint(1, 8), type uniforms, position normals) or injected (``draws=``) for parity tests.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

TYPE_DIST = {'uniform': 0, 'absorbing': 1, 'zeros': 1}
POS_DIST = {'gaussian': 0, 'zero_mean_gaussian': 1}
# len(map_atom_type_aromatic_to_index) / len(map_atom_type_only_to_index), repo/utils/molecule/constants.py:54-80
NUM_TYPES = {'add_aromatic': 13, 'basic': 8}


class SizePrior:
    """``_atom_num_dist.npy`` ({'bounds': [...], 'bins': [(values, probs), ...]}) as flat arrays; ``cdf`` is what numpy's
    legacy ``choice(values, p=probs)`` searches (init_lig.py:27-31)."""

    def __init__(self, bounds, bin_ptr, values, probs):
        self.bounds = np.ascontiguousarray(bounds, dtype=np.float64)
        self.bin_ptr = np.ascontiguousarray(bin_ptr, dtype=np.int32)
        self.values = np.ascontiguousarray(values, dtype=np.int32)
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        assert self.bin_ptr.shape[0] == self.bounds.shape[0] + 2 and self.bin_ptr[-1] == self.values.shape[0]
        cdf = np.empty_like(probs)
        for b in range(self.bin_ptr.shape[0] - 1):
            lo, hi = self.bin_ptr[b], self.bin_ptr[b + 1]
            c = np.cumsum(probs[lo:hi])
            cdf[lo:hi] = c / c[-1]
        self.cdf = cdf
        self._dev = {}

    @classmethod
    def from_npy(cls, path):
        """path = the reference's repo/datasets/transforms/_atom_num_dist.npy (or _linker_num_dist.npy, same format)."""
        d = np.load(path, allow_pickle=True).item()
        return cls.from_table(d)

    @classmethod
    def from_table(cls, d):
        ptr = np.cumsum([0] + [len(b[0]) for b in d['bins']])
        return cls(d['bounds'], ptr, np.concatenate([np.asarray(b[0]) for b in d['bins']]),
                   np.concatenate([np.asarray(b[1], dtype=np.float64) for b in d['bins']]))

    def on(self, device):
        key = str(device)
        if key not in self._dev:
            t = lambda a: torch.from_numpy(a).to(device)
            self._dev[key] = (t(self.bounds), t(self.bin_ptr), t(self.values), t(self.cdf))
        return self._dev[key]


class SizePriorStruct(C.Structure):
    _fields_ = [('bounds', C.c_void_p), ('n_bounds', C.c_int32), ('bin_ptr', C.c_void_p), ('values', C.c_void_p),
                ('cdf', C.c_void_p)]


class BatchSpec(C.Structure):
    _fields_ = [('prot_pos', C.c_void_p), ('prot_element', C.c_void_p), ('prot_backbone', C.c_void_p), ('prot_aa', C.c_void_p),
                ('prot_ptr', C.c_void_p), ('n_pockets', C.c_int32), ('repeat', C.c_int32), ('centre', C.c_void_p),
                ('ctx_pos', C.c_void_p), ('ctx_type', C.c_void_p), ('ctx_ptr', C.c_void_p), ('lig_ptr', C.c_void_p),
                ('pos_noise', C.c_void_p), ('type_u', C.c_void_p), ('num_classes', C.c_int32), ('type_dist', C.c_int32),
                ('pos_dist', C.c_int32),
                ('protein_pos', C.c_void_p), ('protein_atom_feature', C.c_void_p), ('protein_aa_type', C.c_void_p),
                ('protein_element_batch', C.c_void_p), ('protein_translation', C.c_void_p), ('ligand_pos', C.c_void_p),
                ('ligand_atom_type', C.c_void_p), ('ligand_element_batch', C.c_void_p), ('ligand_ctx_flag', C.c_void_p),
                ('ligand_gen_flag', C.c_void_p)]


def _get(d, k, default=None):
    return d.get(k, default) if hasattr(d, 'get') else getattr(d, k, default)


class DeviceBatchBuilder:
    """recipe 'denovo' (all ligand atoms generated) or 'context' (fixed context atoms first, generated atoms after)."""

    def __init__(self, prior, recipe='denovo', type_dist='uniform', mode='add_aromatic', pos_dist='gaussian',
                 num_classes=None):
        if recipe not in ('denovo', 'context'):
            raise ValueError(f'unknown recipe {recipe!r}')
        if type_dist not in TYPE_DIST:
            raise NotImplementedError(f'atom-type distribution {type_dist!r} is not used by the TargetDiff / DiffBP / DiffSBDD '
                                      'test configs and is not implemented')
        if pos_dist not in POS_DIST:
            raise NotImplementedError(f'position distribution {pos_dist!r}')
        if recipe == 'context' and pos_dist != 'gaussian':
            raise NotImplementedError('context tasks use assign_genpos(gaussian)')
        self.prior, self.recipe, self.type_dist, self.pos_dist = prior, recipe, type_dist, pos_dist
        self.num_classes = int(num_classes if num_classes is not None else NUM_TYPES[mode])

    @classmethod
    def from_transform_cfg(cls, transforms, prior, num_classes=None):
        """transforms = config.data.test.transform (list of dicts with 'type', configs/*/test/*.yml)."""
        kw = {'recipe': None}
        for t in transforms:
            name = _get(t, 'type')
            if name in ('featurize_protein_fa', 'merge', 'choose_ctx_gen'):
                continue
            if name == 'remove_ligand':
                kw['recipe'] = 'denovo'
            elif name == 'remove_ligand_gen':
                kw['recipe'] = 'context'
            elif name in ('center_pos', 'center_whole_pos'):
                flag = _get(t, 'center_flag', 'protein') if name == 'center_pos' else 'protein'
                kw['_centre'] = 'context' if flag == 'ligand' else 'denovo'
                if flag == 'ligand' and _get(t, 'mask_flag') != 'ctx_flag':
                    raise NotImplementedError('center_pos(center_flag=ligand) needs mask_flag=ctx_flag')
            elif name in ('assign_molsize', 'assign_gensize'):
                if _get(t, 'distribution', 'prior_distcond') != 'prior_distcond':
                    raise ValueError('only prior_distcond exists in the reference (init_lig.py:240-246)')
            elif name in ('assign_atomtype', 'assign_genatomtype'):
                kw['type_dist'] = _get(t, 'distribution', 'uniform')
                kw['mode'] = _get(t, 'mode', 'add_aromatic')
            elif name in ('assign_molpos', 'assign_genpos'):
                kw['pos_dist'] = _get(t, 'distribution', 'gaussian')
            else:
                raise NotImplementedError(f'transform {name!r} is not part of the sampling batch builder')
        centre = kw.pop('_centre', kw['recipe'])
        if kw['recipe'] is None or centre != kw['recipe']:
            raise NotImplementedError('transform list does not match a shipped sampling recipe (configs/*/test/*.yml)')
        return cls(prior, num_classes=num_classes, **kw)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def build(self, pockets, repeat, device=None, context=None, draws=None, generator=None):
        """pockets: list of raw pockets {'pos' [P,3] f32, 'element' [P] int, 'is_backbone' [P] bool, 'atom_to_aa_type' [P] int}
        (CPU or device tensors); context (recipe 'context'): list of {'pos' [C,3], 'atom_type' [C]} per pocket.
        Every pocket is sampled ``repeat`` times (sample.py:177: config.sampling.num_samples); graph id = sample index.
        draws (optional, parity tests): {'u_size' [S] f64, 'extra' [S] i32, 'type_u' [n,K] f32, 'pos_noise' [n,3] f32}.
        Returns the flat batch dict on the device (+ 'space_size', 'n_lig', 'graph_pocket', 'graph_translation')."""
        L = _lib.lib()
        dev = torch.device(device if device is not None else 'cuda')
        if dev.type != 'cuda':
            raise RuntimeError('DeviceBatchBuilder runs on a CUDA device only (no CPU fallback)')
        n_p = len(pockets)
        if n_p == 0 or repeat <= 0:
            raise ValueError('need at least one pocket and repeat >= 1')
        cat = lambda key, dt: torch.cat([torch.as_tensor(p[key]).reshape(-1, *torch.as_tensor(p[key]).shape[1:]) for p in pockets]).to(dev, dt).contiguous()
        prot_pos = cat('pos', torch.float32)
        prot_el = cat('element', torch.int32)
        prot_bb = cat('is_backbone', torch.uint8)
        prot_aa = cat('atom_to_aa_type', torch.int32)
        counts = [int(torch.as_tensor(p['pos']).shape[0]) for p in pockets]
        if min(counts) < 2:
            raise ValueError('a pocket needs at least two atoms (space size = median of pair distances)')
        prot_ptr = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32, device=dev)
        ctx_pos = ctx_type = ctx_ptr = None
        if self.recipe == 'context':
            if context is None or len(context) != n_p:
                raise ValueError("recipe 'context' needs one context entry per pocket")
            cc = [int(torch.as_tensor(c['pos']).shape[0]) for c in context]
            ctx_pos = torch.cat([torch.as_tensor(c['pos']).reshape(-1, 3) for c in context]).to(dev, torch.float32).contiguous()
            ctx_type = torch.cat([torch.as_tensor(c['atom_type']).reshape(-1) for c in context]).to(dev, torch.int32).contiguous()
            ctx_ptr = torch.tensor(np.concatenate([[0], np.cumsum(cc)]), dtype=torch.int32, device=dev)
            if ctx_pos.numel() == 0:       # keep valid pointers for the C call
                ctx_pos = torch.zeros(1, 3, device=dev)
                ctx_type = torch.zeros(1, dtype=torch.int32, device=dev)
        elif context is not None:
            raise ValueError("recipe 'denovo' takes no context atoms")
        S = n_p * repeat
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            space = torch.empty(n_p, dtype=torch.float32, device=dev)
            centre = torch.empty(n_p, 3, dtype=torch.float32, device=dev)
            _lib.check(L.cbg_pocket_stats_f32(prot_pos.data_ptr(), prot_ptr.data_ptr(), n_p, _lib.ptr(ctx_pos), _lib.ptr(ctx_ptr),
                                              1 if self.recipe == 'context' else 0, space.data_ptr(), centre.data_ptr(), st))
            d = draws or {}
            u_size = (d['u_size'].to(dev, torch.float64) if 'u_size' in d
                      else torch.rand(S, dtype=torch.float64, device=dev, generator=generator)).contiguous()
            extra = None
            if self.recipe == 'context':
                extra = (d['extra'].to(dev, torch.int32) if 'extra' in d
                         else torch.randint(1, 8, (S,), device=dev, generator=generator, dtype=torch.int32)).contiguous()
            bounds, bin_ptr, values, cdf = self.prior.on(dev)
            prior = SizePriorStruct(bounds.data_ptr(), int(bounds.shape[0]), bin_ptr.data_ptr(), values.data_ptr(), cdf.data_ptr())
            n_lig = torch.empty(S, dtype=torch.int32, device=dev)
            lig_ptr = torch.empty(S + 1, dtype=torch.int32, device=dev)
            _lib.check(L.cbg_sample_ligand_sizes(C.byref(prior), space.data_ptr(), n_p, repeat, u_size.data_ptr(),
                                                 _lib.ptr(ctx_ptr), _lib.ptr(extra), n_lig.data_ptr(), lig_ptr.data_ptr(), st))
            n_total = int(lig_ptr[-1].item())            # the one host read-back: output sizes are data dependent
            K = self.num_classes
            pos_noise = (d['pos_noise'].to(dev, torch.float32) if 'pos_noise' in d
                         else torch.randn(n_total, 3, device=dev, generator=generator)).contiguous()
            type_u = None
            if self.type_dist == 'uniform':
                type_u = (d['type_u'].to(dev, torch.float32) if 'type_u' in d
                          else torch.rand(n_total, K, device=dev, generator=generator)).contiguous()
                assert type_u.shape == (n_total, K)
            assert pos_noise.shape == (n_total, 3)
            n_atoms = int(prot_ptr[-1].item()) * repeat
            out = {
                'protein_pos': torch.empty(n_atoms, 3, device=dev), 'protein_atom_feature': torch.empty(n_atoms, 7, device=dev),
                'protein_aa_type': torch.empty(n_atoms, dtype=torch.int64, device=dev),
                'protein_element_batch': torch.empty(n_atoms, dtype=torch.int64, device=dev),
                'protein_translation': torch.empty(n_atoms, 3, device=dev),
                'ligand_pos': torch.empty(n_total, 3, device=dev), 'ligand_atom_type': torch.empty(n_total, dtype=torch.int64, device=dev),
                'ligand_element_batch': torch.empty(n_total, dtype=torch.int64, device=dev),
            }
            ctx_flag = gen_flag = None
            if self.recipe == 'context':
                ctx_flag = torch.empty(n_total, dtype=torch.bool, device=dev)
                gen_flag = torch.empty(n_total, dtype=torch.bool, device=dev)
            spec = BatchSpec(prot_pos.data_ptr(), prot_el.data_ptr(), prot_bb.data_ptr(), prot_aa.data_ptr(), prot_ptr.data_ptr(),
                             n_p, repeat, centre.data_ptr(), _lib.ptr(ctx_pos), _lib.ptr(ctx_type), _lib.ptr(ctx_ptr),
                             lig_ptr.data_ptr(), pos_noise.data_ptr(), _lib.ptr(type_u), K, TYPE_DIST[self.type_dist],
                             POS_DIST[self.pos_dist], out['protein_pos'].data_ptr(), out['protein_atom_feature'].data_ptr(),
                             out['protein_aa_type'].data_ptr(), out['protein_element_batch'].data_ptr(),
                             out['protein_translation'].data_ptr(), out['ligand_pos'].data_ptr(), out['ligand_atom_type'].data_ptr(),
                             out['ligand_element_batch'].data_ptr(), _lib.ptr(ctx_flag), _lib.ptr(gen_flag))
            _lib.check(L.cbg_build_batch_f32(C.byref(spec), st))
        out['protein_lig_flag'] = torch.zeros(n_atoms, dtype=torch.bool, device=dev)
        out['ligand_lig_flag'] = torch.ones(n_total, dtype=torch.bool, device=dev)
        if ctx_flag is not None:
            out['ligand_ctx_flag'], out['ligand_gen_flag'] = ctx_flag, gen_flag
        if self.type_dist == 'zeros':        # AssignMolType('zeros'): [n, K] integer zeros (init_lig.py:389-390,399-400)
            out['ligand_atom_type'] = torch.zeros(n_total, K, dtype=torch.int64, device=dev)
        out['space_size'], out['n_lig'] = space, n_lig
        out['graph_pocket'] = torch.arange(S, device=dev) // repeat
        out['graph_translation'] = centre[out['graph_pocket']]
        return out
