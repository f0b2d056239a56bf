"""B200 drop-in for the reference's ``DiffBP`` model, sampling path.

Mirrors /root/reference repo/models/diffusion/diffbp.py:30-57 (``CoMPredictor`` parameters), :103-130 (constructor,
sub-module names => state-dict keys) and :240-299 (``sample(batch) -> traj``).  Per step ONE C-ABI call
(``cbg_bp_step_f32``): ligand embedding -> kNN -> edge gate -> 9 x (X2H, H2X) -> classifier -> CoM head (its own
gate + 3 x H2X on the final h, from the step's input coordinates) -> fused reverse step (zero-mean noise prediction
+ per-graph CoM shift, score-form VP update, mask-type update).

The pocket is static, so the step-invariant caches of the TargetDiff path (R-cache, static neighbour lists, cached
gates, receptive-field pruning) all apply to the denoiser part unchanged.

Random numbers: ``torch.randn_like`` (positions) then ``torch.rand_like`` over [n_lig] (type change mask) per step,
in the reference's order (diffusion_scheduler.py:158, 486), or injected for parity tests.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from .modules import (GaussianSmearing, H2XAttention, MLP, _NoTorchPath, cfg_get, pack_denoiser_blob)
from .schedulers import CTNVPTables
from .targetdiff import BaseDiffB200, register_model

ABSORBING_STATE = 0      # repo/utils/molecule/constants.py:8


class CoMPredictorB200(_NoTorchPath):
    """Parameter container for CoMPredictor (diffbp.py:30-57)."""

    def __init__(self, cfg):
        super().__init__()
        self.hidden_dim = cfg_get(cfg, 'node_feat_dim', 128)
        self.n_heads = cfg_get(cfg, 'n_heads', 16)
        self.num_r_gaussian = cfg_get(cfg, 'num_r_gaussian', 20)
        self.num_layers = cfg_get(cfg, 'num_layers_com', 3)
        if cfg_get(cfg, 'ew_type', 'global') != 'global' or cfg_get(cfg, 'cutoff_mode', 'knn') != 'knn':
            raise NotImplementedError("CoMPredictorB200 needs ew_type == 'global' and cutoff_mode == 'knn' "
                                      "(the reference's radius branch is broken, diffbp.py:60)")
        self.h2xattentions = nn.ModuleList([H2XAttention(self.hidden_dim, self.n_heads, cfg_get(cfg, 'edge_feat_dim', 4),
                                                         self.num_r_gaussian) for _ in range(self.num_layers)])
        self.dist_emb = nn.Sequential(GaussianSmearing(self.num_r_gaussian),
                                      MLP(self.num_r_gaussian, 1, self.num_r_gaussian * 8))
        self._blob = None
        self._blob_key = None

    def packed_blob(self, device):
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())
        if self._blob is None or key != self._blob_key:
            self._blob = pack_denoiser_blob(dict(self.state_dict()), '', self.num_layers, 1, com_head=True).to(device)
            self._blob_key = key
        return self._blob


class MaskTypeScheduleB200(nn.Module):
    """MaskTypeSchedule (diffusion_scheduler.py:444-450): no parameters, only the change probability of a step."""

    def __init__(self, num_timestep, num_classes, absorbing_state, type='uniform'):
        super().__init__()
        self.num_timestep, self.num_classes = num_timestep, num_classes
        self.absorbing_state, self.schedule_type = absorbing_state, type
        if absorbing_state != 0:
            raise NotImplementedError('the fused reverse step assumes absorbing_state == 0 (constants.py:8)')

    def change_prob(self, t_idx):
        """((T - t) / T).clamp(0, 1) in fp32 like diffusion_scheduler.py:482-484."""
        t = torch.tensor([t_idx], dtype=torch.long)
        return float(((self.num_timestep - t) / self.num_timestep).clamp(max=1., min=0.)[0])


@register_model('diffbp')
class DiffBPB200(BaseDiffB200):
    def __init__(self, cfg):
        super().__init__(cfg)
        gen = cfg.generator
        ps = gen.pos_schedule
        self.pos_scheduler = CTNVPTables(self.num_diffusion_timesteps, beta_start=ps.beta_start,
                                         beta_end=ps.beta_end, type=ps.type)
        self.type_scheduler = MaskTypeScheduleB200(self.num_diffusion_timesteps, num_classes=self.num_classes,
                                                   type=gen.atom_schedule.type, absorbing_state=ABSORBING_STATE)
        self._build_networks(cfg)
        if cfg_get(cfg.encoder, 'cutoff_mode', 'knn') != 'knn':
            raise NotImplementedError('DiffBPB200: the CoM head shares the kNN graph of the denoiser')
        self.com_head = CoMPredictorB200(cfg.encoder)
        self.intersect_reg = cfg.get('intersect_reg', True) if hasattr(cfg, 'get') else True

    def step_coef(self, t_idx):
        ps = self.pos_scheduler
        return _lib.BpCoef(alpha_cumprod=float(ps.host_table('alphas_cumprod')[t_idx]),
                           beta=float(ps.host_table('betas')[t_idx]),
                           nonzero=0.0 if t_idx == 0 else 1.0,
                           change_prob=self.type_scheduler.change_prob(t_idx))

    @torch.no_grad()
    def run_steps(self, state, t_seq, X, Cc, pos_noise=None, type_uniform=None, eps_out=None):
        """Enqueue the reverse steps ``t_seq`` (descending t): one ``cbg_bp_step_f32`` call each.  X / Cc as in
        TargetDiffB200.run_steps (slot t+1 = state entering step t, slot t = its result)."""
        self.check_state(state)
        dev, n_lig, plan = state['device'], state['n_lig'], state['plan']
        com_blob = self.com_head.packed_blob(dev)
        v_scratch = torch.empty(n_lig, dtype=torch.int64, device=dev)
        L = _lib.lib()
        st = _lib.stream_ptr(dev)
        launches0 = L.cbg_launch_count()
        with torch.cuda.device(dev):
            for t_idx in t_seq:
                x_t, c_t = X[t_idx + 1], Cc[t_idx + 1]
                eps = torch.randn_like(x_t) if pos_noise is None else pos_noise[t_idx].to(dev, torch.float32).contiguous()
                uni = (torch.rand((n_lig,), device=dev) if type_uniform is None
                       else type_uniform[t_idx].to(dev, torch.float32).contiguous())
                e_buf = None
                if eps_out is not None:
                    e_buf = eps_out[t_idx] = torch.empty((n_lig, 3), dtype=torch.float32, device=dev)
                coef = self.step_coef(t_idx)
                _lib.check(L.cbg_bp_step_f32(C.byref(plan), com_blob.data_ptr(), self.com_head.num_layers, C.byref(coef),
                                             x_t.data_ptr(), c_t.data_ptr(), eps.data_ptr(), uni.data_ptr(),
                                             X[t_idx].data_ptr(), Cc[t_idx].data_ptr(), v_scratch.data_ptr(),
                                             e_buf.data_ptr() if e_buf is not None else None, None, st))
        self.last_launches = L.cbg_launch_count() - launches0

    @torch.no_grad()
    def sample(self, batch, pos_noise=None, type_uniform=None, num_steps=None, traj_mode='full', eps_out=None):
        """DiffBP.sample (diffbp.py:240-299).  Returns ``traj``: {t: (x_lig, c_lig one-hot, batch_idx_lig)}, keys
        T-1 ... -1, entries >= 0 on the CPU and key -1 on the device like the reference.

        ``pos_noise[t]`` [n_lig,3] / ``type_uniform[t]`` [n_lig] inject the random numbers; ``num_steps`` stops early;
        ``traj_mode='final'`` keeps only traj[0] and traj[-1]; ``eps_out`` (dict) receives eps + eps_com per step."""
        T, K = self.num_diffusion_timesteps, self.num_classes
        state = self.prepare(batch)
        dev, n_lig = state['device'], state['n_lig']
        X = torch.empty((T + 1, n_lig, 3), dtype=torch.float32, device=dev)
        Cc = torch.empty((T + 1, n_lig, K), dtype=torch.float32, device=dev)
        X[T].copy_(state['x_lig'])
        Cc[T].copy_(state['c_lig'])
        t_seq = list(reversed(range(T)))
        if num_steps is not None:
            t_seq = t_seq[:num_steps]
        self.run_steps(state, t_seq, X, Cc, pos_noise, type_uniform, eps_out)
        bl = state['batch_idx_lig']
        t_last = t_seq[-1]
        traj = {}
        bl_cpu = bl.cpu()
        if traj_mode == 'full':
            Xh, Ch = X[t_last + 1:].cpu(), Cc[t_last + 1:].cpu()
            for t in range(t_last, T):
                traj[t] = (Xh[t - t_last], Ch[t - t_last], bl_cpu)
        else:
            traj[t_last] = (X[t_last + 1].cpu(), Cc[t_last + 1].cpu(), bl_cpu)
        traj[t_last - 1] = (X[t_last].clone(), Cc[t_last].clone(), bl)
        return traj
