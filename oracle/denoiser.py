"""CPU restatement of the reference's per-step denoiser (UniTransformer) - fp32 torch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional style: every function takes
the reference's own ``state_dict`` tensors (same key names as the reference module
tree) and follows the reference's as-written formulation.

Reference code followed (``/root/reference``):
  repo/modules/common.py:114-133     GaussianSmearing (20 fixed offsets, coeff)
  repo/modules/common.py:61-68       outer_product (edge_type (x) RBF -> 80 wide)
  repo/modules/common.py:151-171     MLP = Linear -> LayerNorm -> ReLU -> Linear
  repo/modules/common.py:174-180     ShiftedSoftplus
  repo/modules/embs/dist_emb.py:6-14 dist_emb = GaussianSmearing -> MLP(20,1,160)
  repo/modules/attention/x2h_attention.py:43-97   X2HAttention.forward
  repo/modules/attention/h2x_attention.py:34-73   H2XAttention.forward
  repo/modules/e3nn/unitransformer.py:88-99       _build_edge_type
  repo/modules/e3nn/unitransformer.py:102-123     UniTransformer.forward
  repo/modules/e3nn/unitransformer.py:167-186     E3DualAttentionLayer.forward
"""
import math
import torch
import torch.nn.functional as F

from . import graph_ops as G

N_HEADS = 16
HIDDEN = 128


def gaussian_smearing(dist, offset):
    """common.py:115-133: coeff = -0.5/(offset[1]-offset[0])**2; exp(coeff*(d-mu)^2)."""
    coeff = -0.5 / float(offset[1] - offset[0]) ** 2
    return torch.exp(coeff * (dist - offset.view(1, -1)) ** 2)


def mlp(sd, prefix, x):
    """common.py:151-171 with num_layer=2, norm=True, act='relu'."""
    y = F.linear(x, sd[prefix + 'net.0.weight'], sd[prefix + 'net.0.bias'])
    y = F.layer_norm(y, (y.shape[-1],), sd[prefix + 'net.1.weight'], sd[prefix + 'net.1.bias'], 1e-5)
    y = F.relu(y)
    return F.linear(y, sd[prefix + 'net.3.weight'], sd[prefix + 'net.3.bias'])


def build_edge_type(src, dst, lig_flag):
    """unitransformer.py:88-99: 0 lig->lig, 1 lig src/prot dst, 2 prot src/lig dst, 3 prot->prot."""
    n_src = lig_flag[src]
    n_dst = lig_flag[dst]
    t = torch.full_like(src, 3)
    t[n_src & n_dst] = 0
    t[n_src & ~n_dst] = 1
    t[~n_src & n_dst] = 2
    return t


def edge_gate(sd, prefix, x, src, dst):
    """unitransformer.py:109-112: e_w = sigmoid(dist_emb(|x_dst - x_src|)), from the INPUT x."""
    dist = torch.norm(x[dst] - x[src], p=2, dim=-1, keepdim=True)
    g = gaussian_smearing(dist, sd[prefix + 'dist_emb.0.offset'])
    return torch.sigmoid(mlp(sd, prefix + 'dist_emb.1.', g))


def _kv_input(sd, lp, x, h, src, dst, etype):
    rel = x[dst] - x[src]
    dist = torch.norm(rel, p=2, dim=-1, keepdim=True)
    g = gaussian_smearing(dist, sd[lp + 'distance_expansion.offset'])            # [E,20]
    onehot = F.one_hot(etype, 4).to(x.dtype)                                      # [E,4]
    r_feat = (onehot[:, :, None] * g[:, None, :]).reshape(len(src), -1)           # [E,80] idx = t*20+m
    kv = torch.cat([onehot, r_feat, h[dst], h[src]], dim=-1)                      # [E,340]
    return kv, rel


def x2h_attention(sd, lp, x, h, src, dst, etype, e_w):
    """x2h_attention.py:43-97 (ew_net_type='global', out_fc=False)."""
    N = h.shape[0]
    kv, _ = _kv_input(sd, lp, x, h, src, dst, etype)
    k = mlp(sd, lp + 'hk_func.', kv).view(-1, N_HEADS, HIDDEN // N_HEADS)
    v = (mlp(sd, lp + 'hv_func.', kv) * e_w).view(-1, N_HEADS, HIDDEN // N_HEADS)
    q = mlp(sd, lp + 'hq_func.', h).view(-1, N_HEADS, HIDDEN // N_HEADS)
    logits = (q[dst] * k / math.sqrt(k.shape[-1])).sum(-1)                        # [E,16]
    alpha = G.scatter_softmax(logits, dst, dim=0, dim_size=N)
    out = G.scatter_sum(alpha.unsqueeze(-1) * v, dst, dim=0, dim_size=N).view(N, HIDDEN)
    return out + h


def h2x_attention(sd, lp, x, h, src, dst, etype, e_w):
    """h2x_attention.py:34-73 (ew_net_type='global')."""
    N = h.shape[0]
    kv, rel = _kv_input(sd, lp, x, h, src, dst, etype)
    k = mlp(sd, lp + 'xk_func.', kv).view(-1, N_HEADS, HIDDEN // N_HEADS)
    v = mlp(sd, lp + 'xv_func.', kv) * e_w.view(-1, 1)                            # [E,16]
    v = v.unsqueeze(-1) * rel.unsqueeze(1)                                        # [E,16,3]
    q = mlp(sd, lp + 'xq_func.', h).view(-1, N_HEADS, HIDDEN // N_HEADS)
    logits = (q[dst] * k / math.sqrt(k.shape[-1])).sum(-1)
    alpha = G.scatter_softmax(logits, dst, dim=0, dim_size=N)
    out = G.scatter_sum(alpha.unsqueeze(-1) * v, dst, dim=0, dim_size=N)          # [N,16,3]
    return out.mean(1)


def num_layers_in(sd, prefix):
    n = 0
    while (prefix + f'blocks.{n}.x2h_layers.0.hk_func.net.0.weight') in sd:
        n += 1
    return n


def unitransformer_forward(sd, x, h, batch_idx, lig_flag, gen_flag, prefix='denoiser.',
                           k=32, cutoff_mode='knn', r_max=10.0, return_trace=False, dtype=torch.float32):
    """unitransformer.py:102-123 (num_blocks=1, ew_type='global').

    cutoff_mode='radius' is OUR definition (oracle/graph_ops.py) - the reference branch
    raises UnboundLocalError (unitransformer.py:76-77)."""
    x = x.to(dtype)      # dtype=float64 (with a float64 state dict) gives a higher-precision
    h = h.to(dtype)      # reference to measure fp32 rounding sensitivity; graphs come from fp32 x
    ptr = G.graph_ptr_from_batch(batch_idx)
    nbr = G.neighbor_table(x, ptr, k=k, r_max=(r_max if cutoff_mode == 'radius' else None))
    edge_index = G.table_to_edge_index(nbr)
    src, dst = edge_index
    etype = build_edge_type(src, dst, lig_flag.bool())
    e_w = edge_gate(sd, prefix, x, src, dst)
    trace = {'nbr': nbr, 'e_w': e_w, 'src': src, 'dst': dst, 'x': [], 'h': []}
    L = num_layers_in(sd, prefix)
    for l in range(L):
        lp = prefix + f'blocks.{l}.'
        # E3DualAttentionLayer.forward (unitransformer.py:167-186): x2h sees the layer-input
        # x; h2x sees the NEW h and the layer-input x; x moves only where gen_flag.
        h = x2h_attention(sd, lp + 'x2h_layers.0.', x, h, src, dst, etype, e_w)
        dx = h2x_attention(sd, lp + 'h2x_layers.0.', x, h, src, dst, etype, e_w)
        x = x + dx * gen_flag.unsqueeze(-1).to(x.dtype)
        if return_trace:
            trace['x'].append(x.clone())
            trace['h'].append(h.clone())
    c = F.linear(h, sd[prefix + 'classifier.0.weight'], sd[prefix + 'classifier.0.bias'])
    c = F.softplus(c) - math.log(2.0)                                             # common.py:174-180
    c = F.linear(c, sd[prefix + 'classifier.2.weight'], sd[prefix + 'classifier.2.bias'])
    if return_trace:
        return x, h, c, trace
    return x, h, c
