"""Restatement of the third-party graph primitives the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference imports these from packages that are neither vendored nor pinned
(README.MD:57-58 installs unversioned ``pyg pytorch-scatter pytorch-cluster``):

* ``torch_geometric.nn.knn_graph``   call site repo/modules/e3nn/unitransformer.py:79-80
* ``torch_geometric.nn.radius_graph`` call site repo/modules/e3nn/unitransformer.py:76-77
  (unreachable in the reference: ``cut_off`` is unbound on that branch)
* ``torch_scatter.scatter_softmax / scatter_sum``
  call sites repo/modules/attention/x2h_attention.py:86,91, h2x_attention.py:67,71
* ``torch_scatter.scatter_mean`` call sites repo/models/diffusion/diffusion_scheduler.py:123-124

Semantics restated here (published behaviour of pytorch-cluster / pytorch-scatter):

knn_graph(x, k, batch, flow='source_to_target'):
    for every centre i take the min(k, n_g - 1) nearest j != i of the SAME graph by
    squared L2 distance in fp32; emit edge_index = [j ; i] (row 0 = source =
    neighbour, row 1 = target = centre), grouped by centre, nearest first.
    Tie rule (implementation-defined upstream): lower index first.
    Squared distance is evaluated as ((dx*dx + dy*dy) + dz*dz) with individually
    rounded fp32 operations (no fused multiply-add) so that the CUDA neighbour
    search can reproduce the selection bit-for-bit.

radius_graph (OUR definition): the k nearest neighbours as above,
    restricted to those with distance <= r (squared distance <= r*r in fp32).

scatter_sum / scatter_mean / scatter_softmax: segment reductions along dim 0 keyed
    by ``index``; softmax is max-subtracted per segment.
"""
import torch


def pairwise_sqdist_f32(xc: torch.Tensor, xa: torch.Tensor) -> torch.Tensor:
    """[m,3],[n,3] -> [m,n] squared distances with per-op fp32 rounding."""
    d = xc[:, None, :] - xa[None, :, :]
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return (dx * dx + dy * dy) + dz * dz


def neighbor_table(x: torch.Tensor, graph_ptr, k: int = 32, r_max=None, width: int = 32):
    """Fixed-width neighbour table: nbr[i, s] = global index of the s-th nearest
    neighbour of i inside its graph (nearest first, ties -> lower index), -1 padded.

    graph_ptr: sequence of B+1 offsets (graphs are contiguous row ranges).
    r_max: if not None, entries with squared distance > r_max**2 are dropped.
    """
    assert k <= width
    x = x.detach().to(torch.float32).cpu()
    N = x.shape[0]
    nbr = torch.full((N, width), -1, dtype=torch.int64)
    r2 = None if r_max is None else torch.tensor(float(r_max), dtype=torch.float32) ** 2
    for g in range(len(graph_ptr) - 1):
        s, e = int(graph_ptr[g]), int(graph_ptr[g + 1])
        n = e - s
        if n <= 1:
            continue
        d2 = pairwise_sqdist_f32(x[s:e], x[s:e])
        d2.fill_diagonal_(float('inf'))
        kk = min(k, n - 1)
        vals, order = torch.sort(d2, dim=1, stable=True)   # stable => lower index on ties
        sel = order[:, :kk] + s
        if r2 is not None:
            sel = torch.where(vals[:, :kk] <= r2, sel, torch.full_like(sel, -1))
        nbr[s:e, :kk] = sel
    return nbr


def table_to_edge_index(nbr: torch.Tensor) -> torch.Tensor:
    """[N,W] table -> edge_index [2,E] = [src=neighbour ; dst=centre], grouped by centre."""
    N, W = nbr.shape
    dst = torch.arange(N)[:, None].expand(N, W)
    m = nbr >= 0
    return torch.stack([nbr[m], dst[m]], dim=0)


def graph_ptr_from_batch(batch_idx: torch.Tensor):
    """Sorted batch vector -> python list of B+1 offsets."""
    b = batch_idx.detach().cpu()
    assert bool((b[1:] >= b[:-1]).all()), "batch_idx must be sorted"
    B = int(b.max()) + 1 if b.numel() else 0
    counts = torch.bincount(b, minlength=B)
    ptr = torch.zeros(B + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr.tolist()


def knn_graph(x, k, batch=None, loop=False, flow='source_to_target', **_):
    assert flow == 'source_to_target' and not loop
    if batch is None:
        batch = torch.zeros(x.shape[0], dtype=torch.int64)
    ptr = graph_ptr_from_batch(batch)
    width = max(32, k)
    nbr = neighbor_table(x, ptr, k=k, width=width)
    return table_to_edge_index(nbr).to(x.device)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow='source_to_target', **_):
    assert flow == 'source_to_target' and not loop
    if batch is None:
        batch = torch.zeros(x.shape[0], dtype=torch.int64)
    ptr = graph_ptr_from_batch(batch)
    nbr = neighbor_table(x, ptr, k=max_num_neighbors, r_max=r, width=max(32, max_num_neighbors))
    return table_to_edge_index(nbr).to(x.device)


def _expand_index(index, src, dim):
    if index.dim() == src.dim():
        return index
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return res.scatter_add_(dim, _expand_index(index, src, dim), src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    if dim < 0:
        dim += src.dim()
    s = scatter_sum(src, index, dim, dim_size=dim_size)
    ones = torch.ones(index.shape[0], dtype=src.dtype, device=src.device)
    cnt = scatter_sum(ones, index, 0, dim_size=s.shape[dim]).clamp(min=1)
    shape = [1] * s.dim()
    shape[dim] = -1
    return s / cnt.view(shape)


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(dim, _expand_index(index, src, dim), src, reduce='amax', include_self=True)
    return res, None


def scatter_softmax(src, index, dim=0, dim_size=None):
    if dim < 0:
        dim += src.dim()
    idx = _expand_index(index, src, dim)
    mx, _ = scatter_max(src, index, dim, dim_size=dim_size)
    ex = (src - mx.gather(dim, idx)).exp()
    den = scatter_sum(ex, index, dim, dim_size=mx.shape[dim])
    return ex / den.gather(dim, idx)
