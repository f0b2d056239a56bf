"""Restatement of what the reference's ``sample.py`` does with the result of ``model.sample(batch)``
(row f1 of SURVEY.md section 8).  TEST INFRASTRUCTURE (see oracle/__init__.py).

* ``translate``                    sample.py:49-52   ``result[0] += translation`` on the CPU
* ``split_batch_into_samples``     sample.py:16-32   per-graph {pos, type}; one-hot types -> arg-max
  (the atomic-number / aromaticity look-ups of sample.py:29-30 belong to the RDKit reconstruction and are not restated)
* ``reference_results``            sample.py:194-206 ``translate(traj[0], batch.protein_translation[:1])`` then the split:
  ONE vector - the first protein atom's - for the whole batch.  Right for the reference's own batches (``num_samples``
  copies of the same pocket, sample.py:177-183), wrong for batches of different pockets.
* ``per_graph_results``            the same with every graph's own vector (what cbgbench_b200.sample_driver does).
"""
import torch


def translate(result, translation):
    pos = result[0].cpu().clone()
    pos += translation.cpu()
    return [pos] + [result[k + 1] for k in range(len(result) - 1)]


def split_batch_into_samples(result):
    batch_idx = result[-1].cpu()
    if batch_idx.numel() == 0:
        return []
    out = []
    for i in range(int(batch_idx.max()) + 1):
        idx = batch_idx == i
        t = result[1].cpu()[idx]
        if t.dim() == 2:
            t = t.argmax(-1)
        out.append({'pos': result[0].cpu()[idx], 'type': t})
    return out


def reference_results(traj0, batch):
    return split_batch_into_samples(translate(traj0, batch['protein_translation'][:1]))


def per_graph_results(traj0, batch):
    x, c, bl = traj0[0].cpu(), traj0[1], traj0[2].cpu()
    br = batch['protein_element_batch'].cpu()
    tr = batch['protein_translation'].cpu()
    B = int(max(int(bl.max()), int(br.max()))) + 1
    per_graph = torch.zeros(B, 3)
    for g in range(B):
        rows = torch.nonzero(br == g).flatten()
        if rows.numel():
            per_graph[g] = tr[rows[0]]
    return split_batch_into_samples([x + per_graph[bl], c, bl])
