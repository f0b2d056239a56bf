"""the D3FG encoder `IPATransformer` (itatransformer.py:14-145) at hidden 256 / 128.

CPU: the oracle restatement (oracle/ipa.py) against the live-reference fixtures (tests/golden/make_golden_f4.py), the
host module's state-dict contract, the packer.  GPU: csrc/ipa.cu through the C-ABI (cbg_ipa_forward_f32) against the
fixtures and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from cbgbench_b200 import _lib, synthetic
from cbgbench_b200.ipatransformer import IPATransformerB200, pack_ipa_blob
from cbgbench_b200.modules import get_e3_gnn
from helpers import GOLDEN as GOLDEN_DIR, assert_close, rel_err

torch.set_grad_enabled(False)
NAMES = ('eps_pos', 'h', 'o_next', 'R_next', 'c')


def _case(case):
    name, hidden, nl, nc, n_nodes, n_lig, seed, gen_mode = case
    model = IPATransformerB200(synthetic.ipa_config(hidden, nl, nc))
    sd = synthetic.seeded_state_dict(model, seed=synthetic.IPA_WEIGHT_SEED, skip_prefixes=())
    model.load_state_dict(sd, strict=True)
    return model, sd, synthetic.make_ipa_inputs(hidden, n_nodes, n_lig, seed, gen_mode)


def _gold():
    return np.load(os.path.join(GOLDEN_DIR, 'ipa_cases.npz'))


@pytest.mark.parametrize('case', synthetic.IPA_CASES, ids=[c[0] for c in synthetic.IPA_CASES])
def test_oracle_matches_reference_fixtures(case):
    from oracle import ipa as OI
    model, sd, inp = _case(case)
    gold = _gold()
    for a, nm in zip(OI.ipatransformer_forward(sd, *inp), NAMES):
        assert rel_err(a, torch.from_numpy(gold[f'{case[0]}/{nm}'])) < 1e-5, nm


def test_state_dict_keys_and_factory():
    with open(os.path.join(GOLDEN_DIR, 'ipa_state_keys.json')) as f:
        want = json.load(f)
    model = IPATransformerB200(synthetic.ipa_config(256, 3, 8))
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert list(got.keys()) == list(want.keys()) and got == want
    for spelling in ('ipatransformer', 'itatransformer'):      # the shipped config uses the second (d3fg_fg.yml:4)
        cfg = synthetic.ipa_config(256, 2, 8)
        cfg['type'] = spelling
        assert isinstance(get_e3_gnn(cfg), IPATransformerB200)
    with pytest.raises(NotImplementedError):
        IPATransformerB200(synthetic.ipa_config(192, 2, 8))
    with pytest.raises(NotImplementedError):
        IPATransformerB200(synthetic.ipa_config(256, 2, 8, cutoff_mode='radius'))
    with pytest.raises(RuntimeError):
        m = IPATransformerB200(synthetic.ipa_config(128, 1, 8))
        x, o, h, b, lig, gen = synthetic.make_ipa_inputs(128, [10], [2], 1)
        m(x, o, h, b, lig, gen)                                  # CPU tensors: no fallback


def test_packer_places_reference_weights():
    model, sd, _ = _case(synthetic.IPA_CASES[0])
    H, L = 256, _lib.lib()
    blob = pack_ipa_blob(sd, H, 3, 1, 8)
    g0 = _lib.blob_layout()['global_floats']
    hf, lf = L.cbg_ipa_head_floats(H), L.cbg_ipa_layer_floats(H)
    assert blob.numel() == g0 + hf + 3 * lf
    name = lambda fn, i: fn(i).decode()
    layer = {name(L.cbg_ipa_layer_field_name, f): (L.cbg_ipa_layer_field_offset(H, f), L.cbg_ipa_layer_field_size(H, f))
             for f in range(L.cbg_ipa_layer_fields())}
    base = g0 + hf + 2 * lf
    w0 = sd['blocks.2.x2h_layers.0.hv_func.net.0.weight']                  # [256, 4 + 80 + 512]
    o, n = layer['V_WRF']
    assert torch.equal(blob[base + o: base + o + n].view(4, 20, H)[1, 6], w0[:, 4 + 20 + 6])
    o, n = layer['NODE_WT']
    wt = blob[base + o: base + o + n].view(H, 5 * H)
    assert torch.equal(wt[9, H:2 * H], w0[:, 84 + H + 9])                  # Pj_v plane = h_src block of the value MLP
    assert torch.equal(wt[9, 3 * H:4 * H], w0[:, 84 + 9])                  # Pi_v plane = h_dst block
    o, n = layer['Q_W1T']
    assert torch.allclose(blob[base + o: base + o + n].view(H, H).t() * 4.0, sd['blocks.2.x2h_layers.0.hq_func.net.3.weight'])


@pytest.mark.gpu
@pytest.mark.parametrize('case', synthetic.IPA_CASES, ids=[c[0] for c in synthetic.IPA_CASES])
def test_cuda_matches_reference_fixtures_and_oracle(case):
    from oracle import ipa as OI
    dev = torch.device('cuda:0')
    model, sd, inp = _case(case)
    model = model.to(dev)
    gold = _gold()
    got = [t.cpu() for t in model(*[t.to(dev) for t in inp])]
    want = OI.ipatransformer_forward(sd, *inp)
    x, o, h, b, lig, gen = inp
    for a, w, nm in zip(got, want, NAMES):
        g = torch.from_numpy(gold[f'{case[0]}/{nm}'])
        assert rel_err(a, g) < 1e-4, (nm, rel_err(a, g))
        assert_close(a, g, what=f'{case[0]}/{nm} vs reference')
        assert_close(a, w, what=f'{case[0]}/{nm} vs oracle')
    assert torch.equal(got[2][~gen], o[~gen])                    # orientation moves only where gen_flag
    assert torch.equal(got[0][~gen], torch.zeros_like(got[0][~gen]))


@pytest.mark.gpu
def test_cuda_graphs_are_independent_and_shared_blocks():
    """A graph's rows do not depend on the other graphs of the batch (bit for bit), and num_blocks = 2 (shared blocks,
    itatransformer.py:115-125) equals the oracle's two passes."""
    from oracle import ipa as OI
    dev = torch.device('cuda:0')
    name, hidden, nl, nc, n_nodes, n_lig, seed, gen_mode = synthetic.IPA_CASES[1]
    model = IPATransformerB200(synthetic.ipa_config(hidden, nl, nc, num_blocks=2))
    sd = synthetic.seeded_state_dict(model, seed=3, skip_prefixes=())
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    inp = synthetic.make_ipa_inputs(hidden, n_nodes, n_lig, seed, gen_mode)
    got = [t.cpu() for t in model(*[t.to(dev) for t in inp])]
    want = OI.ipatransformer_forward(sd, *inp, num_blocks=2)
    for a, w, nm in zip(got, want, NAMES):
        assert rel_err(a, w) < 1e-4, (nm, rel_err(a, w))
    b = inp[3]
    m = b == 1
    sub = [t[m] for t in inp]
    sub[3] = torch.zeros(int(m.sum()), dtype=torch.long)
    alone = [t.cpu() for t in model(*[t.to(dev) for t in sub])]
    for a, full in zip(alone, got):
        assert torch.equal(a, full[m])
