"""Multi-GPU data path: independent pockets shard embarrassingly, one collective at the end.

The reference is single-process / single-GPU (sample.py:107,155); graphs never interact
(kNN is per graph, every reduction keys on within-graph destinations), so each rank samples
its own graphs for all T steps with ZERO communication and the final coordinates / atom types
(``traj[0]``, what sample.py:194-206 consumes) are collected with a single all-gather over
NCCL (NVLink 5 / NVSwitch).  SURVEY.md section 8(e).

Host logic only (no CUDA here): works with any ``torch.distributed`` backend, which is how the
world_size-2 gloo tests on CPU exercise it.
"""
import torch
import torch.distributed as dist

BATCH_LIG_KEYS = ('ligand_pos', 'ligand_atom_type', 'ligand_lig_flag', 'ligand_gen_flag', 'ligand_ctx_flag',
                  'ligand_element_batch')
BATCH_REC_KEYS = ('protein_pos', 'protein_atom_feature', 'protein_aa_type', 'protein_lig_flag',
                  'protein_gen_flag', 'protein_translation', 'protein_element_batch')
# 'protein_translation' has the reference's layout: one row per PROTEIN ATOM ([n_rec, 3], center_pos transform,
# translation.py:11-24); the per-graph form used by the driver is 'graph_translation' [B, 3]


def graph_sizes(batch):
    """Atoms per graph (protein + ligand), int64 [B]."""
    bl, br = batch['ligand_element_batch'], batch['protein_element_batch']
    B = int(max(int(bl.max()) if bl.numel() else -1, int(br.max()) if br.numel() else -1)) + 1
    return torch.bincount(bl, minlength=B) + torch.bincount(br, minlength=B)


def assign_graphs(sizes, world_size):
    """Greedy longest-processing-time partition of graphs over ranks, balanced by atom count
    (work per graph is proportional to its atoms: fixed 32 in-edges per atom).
    Returns a list (rank -> sorted list of graph ids).  Deterministic."""
    sizes = [int(s) for s in sizes]
    order = sorted(range(len(sizes)), key=lambda g: (-sizes[g], g))
    load = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for g in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        parts[r].append(g)
        load[r] += sizes[g]
    return [sorted(p) for p in parts]


def take_graphs(batch, graph_ids):
    """Sub-batch holding ``graph_ids`` (ascending), graph ids renumbered 0..len-1."""
    gids = torch.as_tensor(graph_ids, dtype=torch.long)
    B = int(max(batch['ligand_element_batch'].max(), batch['protein_element_batch'].max())) + 1
    remap = torch.full((B,), -1, dtype=torch.long)
    remap[gids] = torch.arange(len(graph_ids))
    out = {}
    for keys, bkey in ((BATCH_LIG_KEYS, 'ligand_element_batch'), (BATCH_REC_KEYS, 'protein_element_batch')):
        b = batch[bkey].cpu()
        keep = remap[b] >= 0
        for k in keys:
            if k in batch:
                out[k] = batch[k][keep.to(batch[k].device)]
        out[bkey] = remap[b[keep]].to(batch[bkey].device)
    if 'graph_translation' in batch:
        out['graph_translation'] = batch['graph_translation'][gids.to(batch['graph_translation'].device)]
    return out


def graph_translation(batch):
    """Per-graph translation [B, 3] that undoes the centring transform, or None.

    Accepts 'graph_translation' [B, 3] (DeviceBatchBuilder) or the reference's per-protein-atom
    'protein_translation' [n_rec, 3] (every atom of a graph carries its graph's vector: the first protein row of each
    graph is taken).  The reference itself adds ``protein_translation[:1]`` to the whole batch (sample.py:198-199),
    which is the same thing for its batches of copies of ONE pocket and wrong for mixed pockets."""
    if 'graph_translation' in batch and batch['graph_translation'] is not None:
        return batch['graph_translation']
    tr = batch.get('protein_translation') if hasattr(batch, 'get') else None
    if tr is None:
        return None
    br = batch['protein_element_batch']
    B = int(max(int(batch['ligand_element_batch'].max()), int(br.max()))) + 1
    if tr.shape[0] != br.shape[0]:
        raise ValueError(f"protein_translation has {tr.shape[0]} rows, expected one per protein atom ({br.shape[0]}); "
                         "use 'graph_translation' for a per-graph tensor")
    first = torch.full((B,), br.shape[0], dtype=torch.long, device=br.device)
    first = first.scatter_reduce(0, br, torch.arange(br.shape[0], device=br.device), reduce='amin')
    out = torch.zeros((B, 3), dtype=tr.dtype, device=tr.device)
    has = first < br.shape[0]                       # graphs without protein atoms keep a zero translation
    out[has.to(tr.device)] = tr[first[has].to(tr.device)]
    return out


def gather_final(x_lig, v_lig, graph_id_global, group=None, counts=None):
    """The ONE collective of the data path: all-gather every rank's final ligand coordinates,
    atom types and (global) graph ids, padded to the largest shard.

    x_lig [n,3] f32, v_lig [n] i64, graph_id_global [n] i64 (all on the backend's device).
    ``counts`` = ligand atoms per rank; every rank can derive it from the deterministic
    partition (sample_sharded does), which keeps the path at exactly one collective.  When it
    is None an extra 8-byte size exchange is done first.
    Returns (x [sum n,3], v [sum n], graph_id [sum n]) ordered by global graph id."""
    world = dist.get_world_size(group)
    if counts is None:
        n = torch.tensor([x_lig.shape[0]], dtype=torch.int64, device=x_lig.device)
        cl = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(cl, n, group=group)
        counts = [int(c) for c in cl]
    counts = [int(c) for c in counts]
    assert counts[dist.get_rank(group)] == x_lig.shape[0]
    m = max(counts) if counts else 0
    packed = torch.zeros((m, 5), dtype=torch.float32, device=x_lig.device)
    packed[: x_lig.shape[0], 0:3] = x_lig
    if x_lig.shape[0]:
        assert int(graph_id_global.max()) < (1 << 24) and int(v_lig.max()) < (1 << 24), 'ids must be exact in fp32'
    packed[: x_lig.shape[0], 3] = v_lig.to(torch.float32)        # class ids < 2^24: exact in fp32
    packed[: x_lig.shape[0], 4] = graph_id_global.to(torch.float32)
    bufs = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(bufs, packed, group=group)            # the single payload collective
    parts = [b[:c] for b, c in zip(bufs, counts)]
    allp = torch.cat(parts, 0)
    order = torch.sort(allp[:, 4], stable=True).indices
    allp = allp[order]
    return allp[:, 0:3].contiguous(), allp[:, 3].round().to(torch.int64), allp[:, 4].round().to(torch.int64)


def sample_sharded(sample_fn, batch, group=None):
    """Shard ``batch`` over the ranks of ``group``, run ``sample_fn(sub_batch) -> (x_lig, v_lig)``
    (final ligand coordinates [n,3] and integer atom types [n] of the sub-batch) on every rank,
    gather.  Every rank returns the full result ordered by global graph id."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = assign_graphs(graph_sizes(batch).tolist(), world)
    mine = parts[rank]
    sub = take_graphs(batch, mine)
    if len(mine):
        x, v = sample_fn(sub)
        gid = torch.as_tensor(mine, dtype=torch.long, device=x.device)[sub['ligand_element_batch'].to(x.device)]
    else:
        dev = batch['ligand_pos'].device
        x = torch.zeros((0, 3), dtype=torch.float32, device=dev)
        v = torch.zeros((0,), dtype=torch.int64, device=dev)
        gid = torch.zeros((0,), dtype=torch.int64, device=dev)
    lig_per_graph = torch.bincount(batch['ligand_element_batch'].cpu(), minlength=len(graph_sizes(batch)))
    counts = [int(lig_per_graph[torch.as_tensor(p, dtype=torch.long)].sum()) if len(p) else 0 for p in parts]
    return gather_final(x, v, gid, group=group, counts=counts)
