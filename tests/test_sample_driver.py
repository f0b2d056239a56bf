"""Row f1: the sample.py-style driver against the oracle restatement of sample.py:16-32, 49-52, 194-206 (split,
translate, the ``[:1]`` quirk of SURVEY.md A12), fed with a torch.save()d batch list that carries the reference's keys
(per-protein-atom ``protein_translation``)."""
import numpy as np
import pytest
import torch

from cbgbench_b200 import sample_driver, sharding, synthetic
from cbgbench_b200.targetdiff import TargetDiffB200

torch.set_grad_enabled(False)


def _batch_with_translation(n_prot, n_lig, seed, same_pocket):
    """Synthetic batch + the centring transform's per-atom translation (translation.py:11-24: every protein atom of a
    graph carries the graph's vector).  same_pocket: the reference's own usage - copies of ONE pocket, different ligands."""
    b = synthetic.make_batch(n_prot, n_lig, seed=seed)
    rs = np.random.RandomState(seed + 1)
    B = len(n_prot)
    centres = torch.from_numpy(rs.normal(0, 20.0, size=(B, 3)).astype(np.float32))
    if same_pocket:
        centres[:] = centres[0]
        n0 = n_prot[0]
        for g in range(1, B):           # same pocket coordinates / features in every graph
            m0, mg = b['protein_element_batch'] == 0, b['protein_element_batch'] == g
            for k in ('protein_pos', 'protein_atom_feature', 'protein_aa_type'):
                b[k][mg] = b[k][m0][:n0]
    b['protein_translation'] = centres[b['protein_element_batch']]
    return b


def test_graph_translation_accepts_both_layouts():
    b = _batch_with_translation([5, 3, 4], [2, 2, 1], seed=3, same_pocket=False)
    tr = sharding.graph_translation(b)
    assert tr.shape == (3, 3)
    for g in range(3):
        assert torch.equal(tr[g], b['protein_translation'][b['protein_element_batch'] == g][0])
    b2 = dict(b, graph_translation=tr + 1.0)
    assert torch.equal(sharding.graph_translation(b2), tr + 1.0)            # the per-graph key wins
    sub = sharding.take_graphs(b, [0, 2])
    assert sub['protein_translation'].shape[0] == sub['protein_pos'].shape[0]
    assert torch.equal(sharding.graph_translation(sub), tr[[0, 2]])
    bad = dict(b, protein_translation=tr)                                    # per-graph tensor under the per-atom key
    with pytest.raises(ValueError):
        sharding.graph_translation(bad)


@pytest.mark.gpu
@pytest.mark.parametrize('same_pocket', [True, False], ids=['copies_of_one_pocket', 'mixed_pockets'])
def test_driver_results_match_oracle_sample_loop(tmp_path, same_pocket):
    from oracle import sample_loop as OS
    T, layers, seed = 4, 2, 11
    n_prot = [60, 60, 60] if same_pocket else [60, 45, 30]
    batches = [_batch_with_translation(n_prot, [8, 5, 11], seed=21, same_pocket=same_pocket),
               _batch_with_translation(n_prot[:2], [6, 9], seed=22, same_pocket=same_pocket)]
    path = str(tmp_path / 'batches.pt')
    torch.save(batches, path)
    got = sample_driver.main(['--batches', path, '--steps', str(T), '--layers', str(layers), '--seed', str(seed)])
    # the same model / generator state outside the driver, post-processed by the oracle's sample.py restatement
    dev = torch.device('cuda:0')
    model = TargetDiffB200(synthetic.targetdiff_config(num_steps=T, num_layers=layers))
    model.load_state_dict(synthetic.seeded_state_dict(model, seed=0), strict=True)
    model = model.to(dev).eval()
    torch.manual_seed(seed)
    want_ref, want_graph = [], []
    for b in batches:
        traj0 = model.sample(b, traj_mode='final')[0]
        want_ref.extend(OS.reference_results(traj0, b))
        want_graph.extend(OS.per_graph_results(traj0, b))
    assert len(got) == len(want_graph) == 5
    for g, wg, wr in zip(got, want_graph, want_ref):
        assert torch.equal(g['v'], wg['type'])
        assert torch.equal(g['pos'], wg['pos'])                 # per-graph translate-back
        if same_pocket:
            assert torch.equal(g['pos'], wr['pos'])             # ... which IS the reference's [:1] form on its own batches
    if not same_pocket:                                          # and differs from it on mixed pockets (A12)
        assert not torch.equal(got[1]['pos'], want_ref[1]['pos'])
