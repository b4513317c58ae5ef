"""Graph-captured, data-parallel training loop on flat buffers.

This is the B200-first counterpart of ``solve_model`` / ``training`` (reference training/training.py:71-243).  The
unchanged reference loop is launch- and sync-bound (DataLoader workers, H2D copies, three ``.item()`` calls and
a host round trip of the Smallify tracker per step); here one training step is

    mask multipliers (+ Smallify EMA)            lfgc_mask_multiplier / lfgc_smallify_ema
    wavelet synthesis -> channels-last grid      lfgc_decode_fwd
    sampler + GT + forward + MSE + backward      lfgc_train_step          (one fused kernel)
    synthesis adjoint + mask gradients           lfgc_decode_bwd / lfgc_mask_param_grad
    [ one NCCL all-reduce of the flat gradient ] torch.distributed (only collective of the path)
    regulariser gradients + Adam                 lfgc_add_l1/l2_grad, lfgc_adam

captured once in a CUDA graph and replayed; the sample stream advances through the device-side step counter, the
learning rate is a device scalar the host decay strategy writes.  All parameters, gradients and Adam moments live
in single flat fp32 buffers; the model's ``nn.Parameter`` objects are views into them, so ``state_dict``,
``save_dropvalues_on_grid`` and the reconstruction path keep working.

Data parallelism: every rank draws ``batch`` samples of the same global Philox stream at its own
offset, the flat gradient is summed with one all-reduce (the loss is pre-scaled by 1/global batch), the
sample-independent regulariser gradients are added after the reduction, and Adam runs redundantly on every rank.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from . import parallel
from ..model.Dropout_Layer import DropoutLayer
from ..model.Feature_Grid_Model import Feature_Grid_Model, _multipliers
from ..model.Smallify_Dropout import SmallifyDropout
from ..model.Straight_Through_Dropout import MaskedWavelet_Straight_Through_Dropout, Straight_Through_Dropout
from ..model.Variational_Dropout_Layer import VariationalDropout


class FastTrainer:

    def __init__(self, model: Feature_Grid_Model, volume: torch.Tensor, batch: int, lr: float = 0.008, seed: int = 0,
                 rank: int = 0, world_size: int = 1, process_group=None, weight_l1: float = 0.0,
                 weight_l2: float = 0.0, use_graph: bool = True, betas=(0.9, 0.999), eps: float = 1e-8,
                 variational: Optional[dict] = None):
        """``variational`` switches the loss to VariationalDropoutLoss (reference Variational_Dropout_Layer.py:33-69,
        training/training.py:116-128): dict(n_voxels, weight_dkl, weight_weights, weight_dkl_multiplier,
        weight_dkl_max=30.0, and either variance_model=<Variance_Model> (dynamic) or log_sigma=<float> (static))."""
        if not volume.is_cuda:
            raise L.LfgcError('FastTrainer needs the volume on the GPU')
        if variational is None and any(isinstance(d, VariationalDropout) and d.d_mask is None for d in model.drop):
            raise L.LfgcError('a model with live variational-dropout masks needs the variational loss: pass '
                              'variational=dict(n_voxels=..., weight_dkl=..., weight_weights=..., '
                              'weight_dkl_multiplier=..., variance_model=... | log_sigma=...)')
        self.var_cfg = dict(variational) if variational is not None else None
        self.var_model = self.var_cfg.get('variance_model') if self.var_cfg else None
        self.model = model
        self.volume = volume.contiguous().float()
        self.batch = int(batch)
        self.rank, self.world = int(rank), int(world_size)
        self.group = process_group
        self.seed = int(seed)
        self.weight_l1, self.weight_l2 = float(weight_l1), float(weight_l2)
        self.betas, self.eps = betas, eps
        self.device = self.volume.device
        model.train()
        self.geom = model.geometry()

        # ---- flat buffers: [coefficients | trainable mask parameters | MLP] --------------------------------------
        self.coeff_params = [p for p in model.feature_grid]
        self.mask_params = []
        for d in model.drop:
            if isinstance(d, DropoutLayer):
                self.mask_params += [p for p in d.parameters()]
        self.mlp_params = model._mlp_params()
        self.var_params = self.var_model._params() if self.var_model is not None else []
        every = self.coeff_params + self.mask_params + self.mlp_params + self.var_params
        self._every_param = every
        # sections start on 16-byte boundaries (the kernels use 128-bit loads where the alignment allows)
        def pad4(n):
            return (n + 3) // 4 * 4
        self.n_coeff_elems = sum(p.numel() for p in self.coeff_params)
        self.n_mask_elems = sum(p.numel() for p in self.mask_params)
        self.mask_off = pad4(self.n_coeff_elems)
        self.mlp_off = self.mask_off + pad4(self.n_mask_elems)
        self.n_mlp_elems = sum(p.numel() for p in self.mlp_params)
        self.var_off = self.mlp_off + pad4(self.n_mlp_elems)
        total = self.var_off + sum(p.numel() for p in self.var_params) if self.var_params \
            else self.mlp_off + self.n_mlp_elems
        with torch.no_grad():
            self.flat_p = torch.zeros(total, device=self.device, dtype=torch.float32)
            self._slices = []
            for group, off in ((self.coeff_params, 0), (self.mask_params, self.mask_off), (self.mlp_params, self.mlp_off),
                               (self.var_params, self.var_off)):
                for p in group:
                    n = p.numel()
                    self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
                    p.data = self.flat_p[off:off + n].view(p.shape)
                    self._slices.append((off, n))
                    off += n
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        # keep the model's own MLP pack coherent with the shared buffer
        def adopt(pack, params, off, count):
            pack.flat = self.flat_p[off:off + count]
            pack._layout = []
            o = 0
            for p in params:
                pack._layout.append((o, p.numel()))
                o += p.numel()
            return pack.flat
        self.mlp_flat = adopt(model._mlp_pack, self.mlp_params, self.mlp_off, self.n_mlp_elems)
        if self.var_params:
            self.var_flat = adopt(self.var_model._pack, self.var_params, self.var_off,
                                  sum(p.numel() for p in self.var_params))
        self._grad_view = {}
        for p, (o, n) in zip(every, self._slices):
            self._grad_view[id(p)] = self.flat_g[o:o + n].view(p.shape)

        self.lr_dev = torch.tensor([lr], device=self.device, dtype=torch.float32)
        self.step_dev = torch.zeros(2, device=self.device, dtype=torch.int32)  # [steps taken, Adam ticket scratch]
        self.loss_sum = torch.zeros(1, device=self.device, dtype=torch.float32)
        self.grid_cl = torch.empty((*self.geom.G, self.geom.Cp), device=self.device, dtype=torch.float32)
        # [grid gradient, channels-last | MLP gradient | loss]: one buffer, so that the grid-step path of a data-parallel
        # run all-reduces ONE message (the adjoint is linear: reducing the grid gradient replaces reducing the coefficient
        # gradients, 0.23 MB instead of 0.45 MB at C16/G15)
        n_grid = self.grid_cl.numel()
        self._p2p = None
        import os
        mode = os.environ.get('LFGC_ALLREDUCE', 'p2p')
        gstep_ok = (os.environ.get('LFGC_GRID_STEP', '1') != '0' and self.var_cfg is None and not self.mask_params
                    and all(s is None for s in model.mask_specs()) and self.weight_l1 == 0.0
                    and ops.grid_step_supported(self.geom))
        # message of the data-parallel step: [grid gradient | K rows of (MLP gradient, loss)]; K > 1 only where the per-sample
        # kernel adds its sums with atomics (peer-sum path): its CTAs are spread over the rows
        self._acc_slices = 8 if (self.world > 1 and mode == 'p2p' and gstep_ok and self.world <= L.MAX_PEERS) else 1
        self._n_red = (n_grid + self._acc_slices * (self.n_mlp_elems + 1) + 3) // 4 * 4
        if self.world > 1 and mode == 'p2p' and gstep_ok and self.world <= L.MAX_PEERS:
            # No collective: the two parity copies of the buffer live in torch symmetric memory, lfgc_peer_sum reads
            # every rank's copy over NVLink behind its own in-kernel barrier (flags at the end of the allocation).
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            buf = symm_mem.empty(2 * self._n_red + 64, dtype=torch.float32, device=self.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, grp)
            bases = [int(a) for a in hdl.buffer_ptrs]
            self._red2 = [buf[:self._n_red], buf[self._n_red:2 * self._n_red]]
            self._p2p = dict(buf=buf, hdl=hdl, rank=int(hdl.rank),
                             peers=[[b + 4 * self._n_red * par for b in bases] for par in (0, 1)],
                             flags=[b + 4 * 2 * self._n_red for b in bases],
                             epoch=torch.zeros(2, device=self.device, dtype=torch.int32),
                             summed=torch.zeros(self._n_red, device=self.device, dtype=torch.float32),
                             ticket=torch.zeros(1, device=self.device, dtype=torch.int32))
            self._p2p['announce'] = ops.peer_announce(self._p2p['flags'], self._p2p['rank'], self._p2p['epoch'],
                                                      self._p2p['ticket'])
            torch.cuda.synchronize()
            dist.barrier(group=grp)          # every rank's flags are zero before anybody's first launch
        else:
            self._red2 = [torch.zeros(self._n_red, device=self.device, dtype=torch.float32)]
        self._side = None
        self._tc_panels = None
        self._par = 0                        # parity of the next step (only the p2p path has two buffers)
        self._set_red(0)
        self.scratch = torch.empty(max(self.geom.decode_scratch_bytes // 4, 4), device=self.device)
        self.workspace = torch.empty(self.geom.backward_workspace_bytes // 4, device=self.device)
        self.steps_done = 0
        self.launches_per_step = None
        self._graphs = {}
        self._use_graph = use_graph
        # staging buffers of the host-fed step (step_host)
        self._in_coords = torch.zeros((self.batch, 3), device=self.device, dtype=torch.float32)
        self._in_targets = torch.zeros(self.batch, device=self.device, dtype=torch.float32)
        self._pipe = None   # lazily built state of step_host_pipelined
        # mask-free models whose per-channel wavelet pyramid fits in shared memory: the whole non-per-sample part of the
        # step (partial reduction + adjoint + Adam + next synthesis) is ONE launch, lfgc_grid_step (LFGC_GRID_STEP=0: off)
        self._gstep = gstep_ok
        # live Smallify masks on every level (test_impl_test / mhd_p_smallify): the multiplier IS the parameter, so the mask
        # kernels disappear -- the betas go straight into the synthesis, the adjoint writes d loss / d beta straight into
        # the flat gradient, the three per-level trackers and the regulariser gradients ride in the Adam
        # kernel (flat EMA buffers): 7 launches per step instead of 18
        self._smallify = (self.var_cfg is None and len(model.drop) > 0
                          and all(isinstance(d, SmallifyDropout) and d.d_mask is None for d in model.drop))
        if self._smallify:
            with torch.no_grad():
                self._ema = torch.cat([d.tracker.EMA.to(self.device).reshape(-1) for d in model.drop]).contiguous()
                self._emavar = torch.cat([d.tracker.EMAVar.to(self.device).reshape(-1) for d in model.drop]).contiguous()
                off = 0
                for d in model.drop:
                    n = d.betas.numel()
                    d.tracker.EMA = self._ema[off:off + n].view(d.betas.shape)
                    d.tracker.EMAVar = self._emavar[off:off + n].view(d.betas.shape)
                    off += n
                assert off == self.n_mask_elems
            self._momentum = float(model.drop[0].tracker.sign_variance_momentum)
        self._gstep_primed = False
        self._gstep_scratch = torch.zeros(max(ops.grid_step_scratch_floats(self.geom), 4), device=self.device) \
            if gstep_ok else None
        self._coeff_offs = []
        off = 0
        for p in self.coeff_params:
            self._coeff_offs.append(off)
            off += p.numel()
        if self.var_cfg is not None:
            cfg = self.var_cfg
            self.var_scale = float(cfg['n_voxels']) / float(self.batch * self.world)   # batch_scale of the reference
            self.w_dkl = torch.full((2,), float(cfg['weight_dkl']), device=self.device, dtype=torch.float64)
            self._s_coords = torch.zeros((self.batch, 3), device=self.device, dtype=torch.float32)
            self._s_gt = torch.zeros(self.batch, device=self.device, dtype=torch.float32)
            self._dlog_sigma = torch.zeros(self.batch, device=self.device, dtype=torch.float32)
            if self.var_model is not None:
                self._log_sigma = torch.zeros(self.batch, device=self.device, dtype=torch.float32)
                self._var_ws = torch.empty(ops.plain_mlp_workspace_floats(self.var_model.width, self.var_model.n_layers),
                                           device=self.device, dtype=torch.float32)
            else:
                self._log_sigma = torch.full((self.batch,), float(cfg['log_sigma']), device=self.device,
                                             dtype=torch.float32)
            self._var_layers = [d for d in model.drop if isinstance(d, VariationalDropout) and d.d_mask is None]
            # the DKL kernel walks the mask section as [log_thetas_i | log_var_i] per layer: check that is what we packed
            sizes, off = [], self.mask_off
            for d in self._var_layers:
                assert d.log_thetas.data_ptr() == self.flat_p.data_ptr() + 4 * off
                assert d.log_var.data_ptr() == self.flat_p.data_ptr() + 4 * (off + d.log_thetas.numel())
                sizes.append(d.log_thetas.numel())
                off += 2 * d.log_thetas.numel()
            self._var_sizes = sizes
            # every level carries a live variational mask (the shipped variational configs): multipliers, their gradients
            # and the noise of ALL levels live in flat buffers, one launch each way (lfgc_variational_multiplier /
            # _param_grad), no per-layer copies or memsets
            self._var_fast = (len(self._var_layers) == len(model.drop) and len(model.drop) > 0
                              and self.n_mask_elems == 2 * sum(sizes))
            if self._var_fast:
                npos = sum(sizes)
                self._vnoise = torch.zeros(npos, device=self.device)
                self._vmult = torch.zeros(npos, device=self.device)
                self._vgmult = torch.zeros(npos, device=self.device)

    # ------------------------------------------------------------------------------------------------------------
    def _set_red(self, par):
        """Views of the [grid gradient | MLP gradient | loss] staging buffer the step with parity ``par`` uses."""
        n_grid = self.grid_cl.numel()
        self.red = self._red2[par % len(self._red2)]
        self.grad_grid = self.red[:n_grid].view_as(self.grid_cl)
        self.red_mlp = self.red[n_grid:n_grid + self._acc_slices * (self.n_mlp_elems + 1)]

    def grad_of(self, p):
        return self._grad_view[id(p)]

    def _allreduce_grads(self):
        """Sum the flat gradient over the ranks (NCCL); returns the tensor that holds the result."""
        torch.distributed.all_reduce(self.flat_g, group=self.group)
        return self.flat_g

    def _fork(self):
        """Side stream for work that is independent of the main chain (the Variance_Model passes of the variational step);
        inside a graph capture the event edges become graph dependencies, so the branches run concurrently on replay."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._side.wait_event(ev)
        return torch.cuda.stream(self._side)

    def _join(self):
        ev = torch.cuda.Event()
        ev.record(self._side)
        torch.cuda.current_stream().wait_event(ev)

    def _step_body(self, host_fed=False):
        model, geom = self.model, self.geom
        in_coords, in_targets = self._in_coords, self._in_targets
        if isinstance(host_fed, tuple):   # ('pipe', b): the pipelined host-fed step reads staging buffer pair b
            in_coords, in_targets = self._pipe['coords'][host_fed[1]], self._pipe['targets'][host_fed[1]]
        if self._gstep:
            return self._step_body_gstep(host_fed, in_coords, in_targets)
        if self._smallify:
            return self._step_body_smallify(host_fed, in_coords, in_targets)
        var_fast = self.var_cfg is not None and getattr(self, '_var_fast', False)
        coeffs = [p.data for p in self.coeff_params]
        if var_fast:
            # fresh xi ~ N(0,1) per mask element and step (Variational_Dropout_Layer.py:105); one draw per level, in level
            # order, so that the generator is consumed exactly like the module path's randn_like calls
            off = 0
            for n_i in self._var_sizes:
                self._vnoise[off:off + n_i].normal_()
                off += n_i
            a_, b_ = self.mask_off, self.mask_off + self.n_mask_elems
            ops.variational_multiplier(self.flat_p[a_:b_], self._vnoise, self._var_sizes, self._vmult, zero_out=self._vgmult)
            specs, mults, off = [], [], 0
            for d, n_i in zip(self.model.drop, self._var_sizes):
                mults.append(self._vmult[off:off + n_i].view(d.log_thetas.shape))
                off += n_i
            auxs = mults
        else:
            specs = model.mask_specs()
            mults, auxs = _multipliers(specs)
        n_global = self.batch * self.world
        vm = self.var_model if self.var_cfg is not None else None
        if self.var_cfg is not None:
            # variational likelihood: the samples are materialised so that the Variance_Model sees the positions; drawing
            # them and the Variance_Model forward do not depend on the synthesis: side branch
            if not host_fed:
                in_coords, in_targets = self._s_coords, self._s_gt
            with self._fork():
                if not host_fed:
                    ops.sample(self.volume.shape, self.batch, seed=self.seed,
                               sample_offset=parallel.sample_stream_offset(0, self.rank, self.batch, self.world),
                               volume=self.volume, step_dev=self.step_dev, step_stride=n_global,
                               out=(None, in_coords, in_targets))
                if vm is not None:
                    ops.plain_mlp_forward(vm.width, vm.n_layers, in_coords, self.var_flat, out=self._log_sigma)
        # the synthesis also clears the grid-gradient accumulator; the fused kernel overwrites loss_sum
        ops.decode_fwd(geom, coeffs, mults, scratch=self.scratch, out=self.grid_cl, also_zero=self.grad_grid)
        if self.var_cfg is None:
            ops.train_step(geom, self.volume, self.batch, self.seed,
                           parallel.sample_stream_offset(0, self.rank, self.batch, self.world),
                           parallel.loss_scale(self.batch, self.world), self.grid_cl,
                           self.mlp_flat, self.grad_grid, self.flat_g[self.mlp_off:], self.loss_sum, self.workspace,
                           step_dev=self.step_dev, step_stride=n_global,
                           coords=in_coords if host_fed else None, targets=in_targets if host_fed else None)
        else:
            self._join()
            ops.train_step(geom, None, self.batch, self.seed, 0, 0.5 * self.var_scale, self.grid_cl, self.mlp_flat,
                           self.grad_grid, self.flat_g[self.mlp_off:self.mlp_off + self.n_mlp_elems], self.loss_sum,
                           self.workspace, coords=in_coords, targets=in_targets, log_sigma=self._log_sigma,
                           dlog_sigma=self._dlog_sigma if vm is not None else None)
            if vm is not None:   # the Variance_Model backward runs beside the synthesis adjoint
                with self._fork():
                    ops.plain_mlp_backward(vm.width, vm.n_layers, in_coords, self._dlog_sigma, self.var_flat,
                                           grad_mlp=self.flat_g[self.var_off:], workspace=self._var_ws)
        if var_fast:
            gviews, off = [], 0
            for d, n_i in zip(self.model.drop, self._var_sizes):
                gviews.append(self._vgmult[off:off + n_i].view(d.log_thetas.shape))
                off += n_i
            ops.decode_bwd(geom, self.grad_grid, coeffs, auxs, [True] * len(gviews), scratch=self.scratch,
                           grad_coeffs=[self.grad_of(p) for p in self.coeff_params], grad_mults=gviews, accumulate=2)
            ops.variational_param_grad(self.flat_p[a_:b_], self._vnoise, self._vgmult, self._var_sizes, self.flat_g[a_:b_])
            gmults = []
        else:
            want = [s is not None and len(s.grad_params) > 0 for s in specs]
            _, gmults = ops.decode_bwd(geom, self.grad_grid, coeffs, auxs, want, scratch=self.scratch,
                                       grad_coeffs=[self.grad_of(p) for p in self.coeff_params])
        for spec, gm in zip(specs, gmults):
            if spec is None or not spec.grad_params:
                continue
            g0, g1 = ops.mask_param_grad(spec.mode, spec.p0.detach(), None if spec.p1 is None else spec.p1.detach(),
                                         spec.noise, gm)
            self.grad_of(spec.grad_params[0]).copy_(g0)
            if len(spec.grad_params) == 2:
                self.grad_of(spec.grad_params[1]).copy_(g1)
        if vm is not None:
            self._join()
        g_red = self._allreduce_grads() if self.world > 1 else self.flat_g
        if self.var_cfg is not None:
            # sample-independent terms of VariationalDropoutLoss, added once after the reduction: KL of the live masks
            # (weight ramped on the device) and weight_weights * sum coeff^2, both times batch_scale
            cfg = self.var_cfg
            if self._var_sizes:
                a = self.mask_off
                b = a + 2 * sum(self._var_sizes)
                ops.variational_dkl_grad(self.flat_p[a:b], g_red[a:b], self._var_sizes, self.w_dkl, self.step_dev,
                                         1.0 + float(cfg['weight_dkl_multiplier']), float(cfg.get('weight_dkl_max', 30.0)),
                                         self.var_scale)
            ww = float(cfg['weight_weights']) * self.var_scale
            if var_fast:   # the weight term rides in the Adam kernel
                ops.adam_reg(self.flat_p, g_red, self.flat_m, self.flat_v, self.lr_dev, self.step_dev,
                             (0, self.n_coeff_elems), ww, (0, 0), 0.0, self.betas[0], self.betas[1], self.eps)
                return
            if ww > 0.0 and self.n_coeff_elems:
                ops.add_l2_grad(g_red[:self.n_coeff_elems], self.flat_p[:self.n_coeff_elems], ww)
        # sample-independent regularisers (SmallifyLoss): added once, after the reduction
        if self.weight_l2 > 0.0 and self.n_coeff_elems:
            ops.add_l2_grad(g_red[:self.n_coeff_elems], self.flat_p[:self.n_coeff_elems], self.weight_l2)
        if self.weight_l1 > 0.0 and self.n_mask_elems:
            a, b = self.mask_off, self.mask_off + self.n_mask_elems
            ops.add_l1_grad(g_red[a:b], self.flat_p[a:b], self.weight_l1)
        ops.adam(self.flat_p, g_red, self.flat_m, self.flat_v, self.lr_dev, self.step_dev, self.betas[0],
                 self.betas[1], self.eps)

    def _step_body_smallify(self, host_fed, in_coords, in_targets):
        """Live Smallify masks (Smallify_Dropout.py:54-61,103-112; SmallifyLoss :22-40): see __init__."""
        geom = self.geom
        n_global = self.batch * self.world
        a, b = self.mask_off, self.mask_off + self.n_mask_elems
        betas = [d.betas.data for d in self.model.drop]          # views into flat_p: the multipliers themselves
        coeffs = [p.data for p in self.coeff_params]
        ops.decode_fwd(geom, coeffs, betas, scratch=self.scratch, out=self.grid_cl, also_zero=self.grad_grid)
        ops.train_step(geom, self.volume, self.batch, self.seed,
                       parallel.sample_stream_offset(0, self.rank, self.batch, self.world),
                       parallel.loss_scale(self.batch, self.world), self.grid_cl,
                       self.mlp_flat, self.grad_grid, self.flat_g[self.mlp_off:], self.loss_sum, self.workspace,
                       step_dev=self.step_dev, step_stride=n_global,
                       coords=in_coords if host_fed else None, targets=in_targets if host_fed else None)
        # d loss / d beta accumulates (atomics) into the mask section of the flat gradient, which the Adam kernel of the
        # previous step left cleared (accumulate = 2: no memsets)
        ops.decode_bwd(geom, self.grad_grid, coeffs, betas, [True] * len(betas), scratch=self.scratch,
                       grad_coeffs=[self.grad_of(p) for p in self.coeff_params],
                       grad_mults=[self.grad_of(d.betas) for d in self.model.drop], accumulate=2)
        g_red = self._allreduce_grads() if self.world > 1 else self.flat_g
        # Adam + SmallifyLoss gradients + the sign-variance tracker (with the betas this step's forward used) + clearing of
        # the mask gradients, one launch
        ops.adam_reg(self.flat_p, g_red, self.flat_m, self.flat_v, self.lr_dev, self.step_dev,
                     (0, self.n_coeff_elems), self.weight_l2, (a, b), self.weight_l1, self.betas[0], self.betas[1], self.eps,
                     ema=self._ema, emavar=self._emavar, momentum=self._momentum, zero_l1_grad=True)

    def _prime_gstep(self):
        """The grid-step path expects the grid of the CURRENT coefficients (and a cleared gradient accumulator) on
        entry: every step leaves them behind for the next one, this provides them for the first."""
        coeffs = [p.data for p in self.coeff_params]
        ops.decode_fwd(self.geom, coeffs, [None] * len(coeffs), scratch=self.scratch, out=self.grid_cl)
        for r in self._red2:
            r.zero_()
        # operand image of the tensor-core kernel: lfgc_grid_step keeps it current from here on
        self._tc_panels = ops.tc_panel_image(self.geom, self.mlp_flat, out=self._tc_panels)
        self._gstep_primed = True
        self._flat_version = self._param_version()

    def _step_body_gstep(self, host_fed, in_coords, in_targets):
        """Two launches per optimiser step: the fused per-sample kernel (partial sums left in the workspace) and
        lfgc_grid_step.  Data parallel (NCCL): the partial reduction stays a launch of its own, [grid gradient | MLP
        gradient] is all-reduced as one message, then lfgc_grid_step."""
        geom = self.geom
        n_global = self.batch * self.world
        pcount = self.n_mlp_elems
        kw = dict(step_dev=self.step_dev, step_stride=n_global, coords=in_coords if host_fed else None,
                  targets=in_targets if host_fed else None)
        offset = parallel.sample_stream_offset(0, self.rank, self.batch, self.world)
        scale = parallel.loss_scale(self.batch, self.world)
        common = dict(grid_cl=self.grid_cl, p=self.flat_p, g=self.flat_g, m=self.flat_m, v=self.flat_v,
                      coeff_offs=self._coeff_offs, mlp_off=self.mlp_off, lr_dev=self.lr_dev, step_dev=self.step_dev,
                      zero_grid=self.grad_grid, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                      weight_l2=self.weight_l2, scratch=self._gstep_scratch, tc_panels=self._tc_panels)
        if self.world == 1:
            ns = ops.train_step_partials(geom, self.volume, self.batch, self.seed, offset, scale, self.grid_cl,
                                         self.mlp_flat, self.grad_grid, self.workspace, tc_panels=self._tc_panels, **kw)
            ops.grid_step(geom, [self.grad_grid], [self.workspace], ns, pcount + 1, pcount, loss_out=self.loss_sum,
                          **common)
            return
        if self._p2p is None:
            ops.train_step(geom, self.volume, self.batch, self.seed, offset, scale, self.grid_cl, self.mlp_flat,
                           self.grad_grid, self.red_mlp[:pcount], self.loss_sum, self.workspace, **kw)
            torch.distributed.all_reduce(self.red, group=self.group)
            ops.grid_step(geom, [self.grad_grid], [self.red_mlp], 1, pcount + 1, pcount, **common)
            return
        # the per-sample kernel adds its MLP-gradient and loss sums straight into the message buffer (cleared one step
        # earlier, like the grid-gradient section): no reduction launch between it and the peer sum
        # ... and its last CTA stores this rank's epoch flags, so they cross NVLink during the launch gap
        ops.train_step_accumulate(geom, self.volume, self.batch, self.seed, offset, scale, self.grid_cl, self.mlp_flat,
                                  self.grad_grid, self.red_mlp, self.workspace, announce=self._p2p['announce'],
                                  n_slices=self._acc_slices, tc_panels=self._tc_panels, **kw)
        # lfgc_peer_sum: this step's buffers of all ranks -> one local sum (barrier inside the kernel); the accumulator it
        # clears is the OTHER parity's (its last readers finished before they announced this epoch)
        P = self._p2p
        par = self._par
        n_grid = self.grid_cl.numel()
        ops.peer_sum(P['peers'][par], P['flags'], P['rank'], P['epoch'], P['summed'], zero=self._red2[par ^ 1],
                     announced=True)
        common['zero_grid'] = None
        ops.grid_step(geom, [P['summed'][:n_grid]], [P['summed'][n_grid:]], self._acc_slices, pcount + 1, pcount, **common)

    def capture(self, host_fed=False):
        """Warm up eagerly (counts launches), then record the step into a CUDA graph; the optimiser state the
        warm-up touched is restored so that training starts from the initial state."""
        state = (self.flat_p.clone(), self.flat_m.clone(), self.flat_v.clone(), self.step_dev.clone())
        trackers = [(d.tracker.EMA.clone(), d.tracker.EMAVar.clone()) for d in self.model.drop
                    if isinstance(d, SmallifyDropout)]
        w_dkl = self.w_dkl.clone() if self.var_cfg is not None else None
        if self._gstep and not self._gstep_primed:
            self._prime_gstep()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):      # an even number of steps: the buffer parity is back where it started
                before = ops.launch_count()
                self._set_red(self._par)
                self._step_body(host_fed)
                self._advance_par()
                self.launches_per_step = ops.launch_count() - before
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        keep = self._par
        for par in ((0, 1) if self._p2p is not None else (0,)):
            graph = None
            if self._use_graph:
                self._par = par
                self._set_red(par)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._step_body(host_fed)
            self._graphs[(host_fed, par)] = graph
        self._par = keep
        self._set_red(keep)
        with torch.no_grad():
            self.flat_p.copy_(state[0])
            self.flat_m.copy_(state[1])
            self.flat_v.copy_(state[2])
            self.step_dev.copy_(state[3])
            if w_dkl is not None:
                self.w_dkl.copy_(w_dkl)
            it = iter(trackers)
            for d in self.model.drop:
                if isinstance(d, SmallifyDropout):
                    ema, var = next(it)
                    d.tracker.EMA.copy_(ema)
                    d.tracker.EMAVar.copy_(var)
        if self._gstep:
            self._prime_gstep()   # the warm-up steps moved the coefficients: decode the restored ones
        torch.cuda.synchronize()
        if self._p2p is not None:
            torch.distributed.barrier(group=self.group)   # nobody starts real steps while a peer still warms up

    def _advance_par(self):
        if self._p2p is not None:
            self._par ^= 1

    def _param_version(self):
        """Changes whenever a torch in-place operation wrote to a parameter or to the flat buffer (the kernels write through
        raw pointers and leave the counters alone)."""
        return self.flat_p._version + sum(p._version for p in self._every_param)

    def refresh(self):
        """Call after changing parameters OUTSIDE the trainer (the grid-step path carries two things derived from them from
        step to step: the decoded grid and the tensor-core operand image).  In-place torch operations on the parameters
        are noticed by themselves (version counters of the parameters and of the flat buffer); raw writes through ``.data``
        are not."""
        if self._gstep:
            self._prime_gstep()

    def _run(self, host_fed):
        if (host_fed, 0) not in self._graphs:
            self.capture(host_fed)
        if self._gstep and self._param_version() != self._flat_version:
            self._prime_gstep()     # somebody wrote to the parameters with torch operations since the last step
        self._set_red(self._par)
        g = self._graphs[(host_fed, self._par)]
        if g is not None:
            g.replay()
        else:
            self._step_body(host_fed)
        self._advance_par()
        self.steps_done += 1

    def step(self):
        """One optimiser step over ``batch`` fresh samples per rank drawn on the device (asynchronous)."""
        self._run(False)

    def step_host(self, coords, targets):
        """One optimiser step on caller-supplied samples -- the reference's DataLoader contract
        (training/training.py:89-109): ``coords`` (batch, 3) normalised positions and ``targets`` (batch,) volume
        values in HOST memory (pinned for asynchronous copies).  Two H2D copies + one graph replay; read the loss with
        ``last_loss()`` (device -> host)."""
        self._in_coords.copy_(coords.view(self.batch, 3), non_blocking=True)
        self._in_targets.copy_(targets.view(self.batch), non_blocking=True)
        self._run(True)

    def step_host_pipelined(self, coords, targets=None):
        """``step_host`` with the transfers taken off the critical path: the H2D copy of step i runs on a copy stream
        into one of two staging buffers while step i-1 computes, and every step's loss lands in pinned host memory.
        Returns the mean squared error of the PREVIOUS call's step (None on the first call); ``flush_host_pipeline()``
        returns the last one.  Same arithmetic as ``step_host``.

        The loop is host-bound (a dozen CUDA API calls per 60 us step), so the calls are kept few: ``coords`` may be ONE
        pinned buffer of 4 * batch floats, [positions (batch x 3) | targets (batch)] (then ``targets`` is None) -- or a
        positions / targets pair that lies back to back in one allocation, which is detected -- for a single H2D copy;
        and the step's loss is written by the kernel straight into pinned host memory (mapped into the device's address
        space), so there is no D2H copy call at all."""
        n = self.batch
        if self._pipe is None:
            stage = [torch.zeros(4 * n, device=self.device) for _ in range(2)]
            self._pipe = dict(
                stage=stage,
                coords=[t[:3 * n].view(n, 3) for t in stage], targets=[t[3 * n:] for t in stage],
                loss=[torch.zeros(1).pin_memory() for _ in range(2)],
                ready=[torch.cuda.Event() for _ in range(2)], done=[torch.cuda.Event() for _ in range(2)],
                stream=torch.cuda.Stream(), i=0)
            keep = self.loss_sum
            for b in range(2):
                self.loss_sum = self._pipe['loss'][b]    # captured as the step's loss destination: host memory, device-visible
                self.capture(('pipe', b))
            self.loss_sum = keep
        P = self._pipe
        b = P['i'] & 1
        cur = torch.cuda.current_stream()
        if targets is None:
            packed = coords.view(-1)
        elif (targets.data_ptr() == coords.data_ptr() + 12 * n and coords.is_contiguous() and targets.is_contiguous()
              and coords.untyped_storage().data_ptr() == targets.untyped_storage().data_ptr()):
            packed = torch.empty(0, dtype=torch.float32).set_(coords.untyped_storage(), coords.storage_offset(), (4 * n,))
        else:
            packed = None
        if P['i'] >= 2:
            P['stream'].wait_event(P['done'][b])          # the step that read this staging buffer has run
        with torch.cuda.stream(P['stream']):
            if packed is not None:
                P['stage'][b].copy_(packed, non_blocking=True)
            else:
                P['coords'][b].copy_(coords.view(n, 3), non_blocking=True)
                P['targets'][b].copy_(targets.view(n), non_blocking=True)
            P['ready'][b].record()
        cur.wait_event(P['ready'][b])
        self._run(('pipe', b))
        if self._p2p is not None:   # peer-sum data parallelism: the loss travels in the summed message (global sum)
            P['loss'][b].copy_((self._summed_loss() / self.world).reshape(1), non_blocking=True)
        P['done'][b].record(cur)
        prev = None
        if P['i'] >= 1:
            P['done'][b ^ 1].synchronize()
            prev = float(P['loss'][b ^ 1]) / self.batch
        P['i'] += 1
        return prev

    def flush_host_pipeline(self):
        """Loss of the last ``step_host_pipelined`` call (waits for it)."""
        P = self._pipe
        if P is None or P['i'] == 0:
            return None
        b = (P['i'] - 1) & 1
        P['done'][b].synchronize()
        return float(P['loss'][b]) / self.batch

    def set_lr(self, lr: float):
        self.lr_dev.fill_(float(lr))

    def _summed_loss(self):
        """Global loss sum of the last peer-summed step: the last float of every row of the message's MLP section."""
        n_grid, row = self.grid_cl.numel(), self.n_mlp_elems + 1
        rows = self._p2p['summed'][n_grid:n_grid + self._acc_slices * row].view(self._acc_slices, row)
        return rows[:, -1].sum()

    def last_loss(self) -> float:
        """Mean squared error of the last step's batch (device -> host read): the local batch, or -- peer-sum data
        parallelism, where the loss travels in the summed message -- the global one."""
        if self._p2p is not None:
            return float(self._summed_loss().item()) / (self.batch * self.world)
        return float(self.loss_sum.item()) / self.batch

    def complete_loss(self) -> float:
        """``complete_loss`` of training/training.py:130-135 for the last step: MSE over the GLOBAL batch plus the
        SmallifyLoss terms (Smallify_Dropout.py:22-40).  The MSE is the step's own (pre-update parameters); the
        regulariser sums are read after the update (one Adam step later than the reference; they carry weights of
        1e-8 in the shipped configs and only feed the plateau test of SmallifyDecayStrategy).  Rank-identical."""
        if self.var_cfg is not None:
            raise L.LfgcError('complete_loss() is defined for the MSE (+ SmallifyLoss) objective only')
        with torch.no_grad():
            if self._p2p is not None:   # already summed over the ranks (identical on every rank)
                t = self._summed_loss().double() / float(self.batch * self.world)
            else:
                t = self.loss_sum.double() / float(self.batch * self.world)
            if self.world > 1 and self._p2p is None:
                t = t.clone()
                torch.distributed.all_reduce(t, group=self.group)
            if self.weight_l1 > 0.0 and self.n_mask_elems:
                t = t + self.weight_l1 * self.flat_p[self.mask_off:self.mask_off + self.n_mask_elems].abs().sum().double()
            if self.weight_l2 > 0.0 and self.n_coeff_elems:
                t = t + self.weight_l2 * (self.flat_p[:self.n_coeff_elems].double() ** 2).sum()
            return float(t.item())


# ---------------------------------------------------------------------------------------------------------------------
# whole training run: the reference's ``training(args)`` flow on the fast loop
# ---------------------------------------------------------------------------------------------------------------------

class LRSchedule:
    """The reference's ``LearningRateDecayStrategy`` pair (training/learning_rate_decay.py:12-57), selected like
    ``create_instance`` (:14-18): ``smallify_decay == 0`` -> NeurcompDecayStrategy (x ``lr_decay`` whenever a pass
    boundary is crossed and (pass + 1) % pass_decay == 0), else SmallifyDecayStrategy (loss plateau: after
    ``smallify_decay`` pass boundaries without a new best loss multiply the rate by ``lr_decay``; once the rate is at or
    below 1e-7 the strategy asks for an early stop).

    ONE object spans both training phases, as in the reference (training/training.py:200,222-237): it stays bound to
    the PHASE-1 optimiser, so in phase 2 it keeps counting (and the plateau flavour can still stop the run) but its
    rate changes no longer reach the optimiser that is stepping.  ``bind(trainer)`` / ``bind(None)`` model that."""

    def __init__(self, args, lr: float):
        self.lr = float(lr)                          # learning rate of the phase-1 optimiser
        self.lr_decay = float(args['lr_decay'])
        self.smallify = int(args.get('smallify_decay', 0) or 0)
        self.epoch_delay = self.smallify if self.smallify else int(args['pass_decay'])
        self.lr_stop = 1e-07
        self.last_loss = None
        self.no_gain_epoch = 0
        self.trainer = None
        self.decays = 0

    def bind(self, trainer):
        self.trainer = trainer

    def _decay(self):
        self.lr *= self.lr_decay
        self.decays += 1
        if self.trainer is not None:
            self.trainer.set_lr(self.lr)

    def update(self, prior_passes: int, cur_passes: float, loss_fn=None) -> bool:
        """Call after every optimiser step; ``loss_fn()`` returns the step's complete loss (only evaluated at a pass
        boundary of the plateau strategy: one device->host read per volume pass).  True = stop training."""
        if prior_passes == int(cur_passes):
            return False
        if not self.smallify:
            if (int(cur_passes) + 1) % self.epoch_delay == 0:
                self._decay()
            return False
        loss = float(loss_fn())
        if self.last_loss is None or loss < self.last_loss:
            self.last_loss = loss
            self.no_gain_epoch = 0
        else:
            self.no_gain_epoch += 1
        if self.no_gain_epoch == self.epoch_delay:
            if self.lr > self.lr_stop:
                self._decay()
            else:
                return True
            self.no_gain_epoch = 0
        return False


def _regulariser_weights(args):
    drop = args.get('drop_type') or ''
    if drop and 'variational' not in drop:
        return float(args['lambda_drop_loss']), float(args['lambda_weight_loss'])
    return 0.0, 0.0


def epoch_schedule(n_voxels: int, batch_size: int, sample_size: int, max_pass: float):
    """The reference's pass accounting as a generator of (step index, prior_passes, volume_passes, last) tuples
    (training/training.py:76-114,178): the ``while int(volume_passes) + 1 < max_pass`` test is evaluated only at
    DataLoader-epoch boundaries (one epoch = ceil(n_voxels / batch_size) batches = exactly ``sample_size`` volume passes,
    because the LAST batch of an epoch is the partial one the DataLoader yields with drop_last=False and the reference
    counts the samples it actually saw, :108); inside an epoch the loop leaves only through
    ``int(volume_passes) >= max_pass``.  The caller may stop early (plateau strategy) by closing the generator.

    The fast loop's batch is fixed (the step is a captured graph), so the one partial step per epoch trains on a full
    batch here (mhd_p: 32768 instead of 12272 samples once every 8097 steps); the ACCOUNTING follows the reference, which
    is what fixes the number of optimiser steps and the learning-rate decay points (30365 steps for the 60-pass mhd_p
    run, tests/golden/psnr_configs.json)."""
    n_voxels, batch_size = int(n_voxels), int(batch_size)
    steps_per_epoch = -(-n_voxels // batch_size)
    last_items = n_voxels - (steps_per_epoch - 1) * batch_size
    seen, passes, step = 0.0, 0.0, 0
    while int(passes) + 1 < max_pass:
        for i in range(steps_per_epoch):
            prior = int(seen / n_voxels)
            seen += float((last_items if i == steps_per_epoch - 1 else batch_size) * sample_size)
            passes = seen / n_voxels
            step += 1
            done = int(passes) >= max_pass
            yield step, prior, passes, done
            if done:
                break


def make_trainer(model, volume, n_voxels, args, lr, seed=0, rank=0, world=1, group=None, regularise=True,
                 variational_sched_check=None):
    """FastTrainer for one phase of the reference's ``training(args)`` (training/training.py:184-237): the objective the
    ``args`` dict selects (MSE, MSE + SmallifyLoss, or VariationalDropoutLoss with the static / dynamic variance), the
    GLOBAL batch ``batch_size * sample_size`` split over ``world`` ranks (strong scaling: the same optimisation problem)."""
    batch = int(args['batch_size']) * int(args['sample_size'])
    if batch % world != 0:
        raise L.LfgcError('batch_size * sample_size = %d is not divisible by the world size %d' % (batch, world))
    w1, w2 = _regulariser_weights(args) if regularise else (0.0, 0.0)
    variational = None
    drop = args.get('drop_type') or ''
    if regularise and 'variational' in drop:
        # reference training/training.py:80-84,116-128,205-209: VariationalDropoutLoss, plus a fresh Variance_Model
        # joining the optimiser for the 'dynamic' flavour, else the constant args['variational_sigma']
        from ..model.Variational_Dropout_Layer import Variance_Model
        variational = dict(n_voxels=n_voxels, weight_dkl=float(args['lambda_drop_loss']),
                           weight_weights=float(args['lambda_weight_loss']),
                           weight_dkl_multiplier=float(args['weight_dkl_multiplier']))
        if 'dynamic' in drop:
            variational['variance_model'] = Variance_Model().to(volume.device).train()
        else:
            variational['log_sigma'] = float(args['variational_sigma'])
        if variational_sched_check is not None and variational_sched_check.smallify:
            raise L.LfgcError('smallify_decay != 0 (loss-plateau schedule) is not implemented for the variational loss '
                              'in the fast loop; use the module path (training/training.py) for that combination')
    return FastTrainer(model, volume, batch // world, lr=lr, seed=seed, rank=rank, world_size=world,
                       process_group=group, weight_l1=w1, weight_l2=w2, variational=variational)


def solve_phase(model, volume, n_voxels, args, max_pass, lr, sched: Optional[LRSchedule] = None, bind: bool = True,
                seed=0, rank=0, world=1, group=None, regularise=True, verbose=False):
    """``solve_model`` (training/training.py:71-181) on the graph-captured step.  Returns (trainer, stopped_early)."""
    trainer = make_trainer(model, volume, n_voxels, args, lr, seed=seed, rank=rank, world=world, group=group,
                           regularise=regularise, variational_sched_check=sched)
    if sched is not None:
        sched.bind(trainer if bind else None)
    stopped = False
    for step, prior, passes, last in epoch_schedule(n_voxels, int(args['batch_size']), int(args['sample_size']),
                                                    max_pass):
        trainer.step()
        if sched is not None and sched.update(prior, passes, trainer.complete_loss):
            stopped = True
            break
        if verbose and trainer.steps_done % 500 == 0:
            print('pass %.3f / %.1f  mse %.6f' % (passes, max_pass, trainer.last_loss()))
    torch.cuda.synchronize()
    return trainer, stopped


def train_volume(args: dict, volume: Optional[torch.Tensor] = None, seed: int = 0, verbose: bool = False, rank: int = 0,
                 world: int = 1, group=None):
    """Two-phase training + evaluation with the reference's schedule (training/training.py:184-243): 2/3 of the
    passes with masks and regularisers, bake the masks, 1/3 fine-tuning at lr/10, strip the mask layers, reconstruct
    the volume and report PSNR / compression ratio."""
    from ..data.IndexDataset import IndexDataset, get_tensor
    from ..model.model_utils import setup_model
    from ..visualization.OutputToVTK import tiled_net_out

    if volume is None:
        volume = get_tensor(args['data'])
    dataset = IndexDataset(volume, args['sample_size'])
    device = torch.device('cuda')
    volume_dev = volume.to(device)
    model = setup_model(args['d_in'], args['n_hidden_size'], args['d_out'], args['n_layers'], args['embedding_type'],
                        args['n_embedding_freq'], args['drop_type'], args['drop_momentum'], args['drop_threshold'],
                        args['wavelet_filter'], args['grid_features'], args['grid_size'], args.get('checkpoint_path', ''))
    model.to(device)
    model.train()
    sched = LRSchedule(args, args['lr'])
    t1, _ = solve_phase(model, volume_dev, dataset.n_voxels, args, args['max_pass'] * (2.0 / 3.0), args['lr'], sched,
                        bind=True, seed=seed, rank=rank, world=world, group=group, verbose=verbose)
    zeros = model.save_dropvalues_on_grid(device)
    # phase 2: plain MSE, lr / 10; the strategy object stays bound to the first optimiser (reference
    # training/training.py:234-237), so its decays do not reach this phase -- but its early stop does
    t2, _ = solve_phase(model, volume_dev, dataset.n_voxels, args, args['max_pass'] * (1.0 / 3.0), args['lr'] / 10.0,
                        sched, bind=False, seed=seed + 1, rank=rank, world=world, group=group, regularise=False,
                        verbose=verbose)
    model.remove_drop_layers(device)
    psnr, l1, mse, rmse = tiled_net_out(dataset, model, True, gt_vol=volume_dev, evaluate=True, write_vols=False)
    n_params = sum(p.numel() for n, p in model.named_parameters() if 'drop' not in n)
    ratio = dataset.n_voxels / (n_params - float(zeros))
    return dict(psnr=psnr, l1_diff=l1, mse=mse, rmse=rmse, num_parameters=n_params, num_zeros=float(zeros),
                compression_ratio=ratio, steps=t1.steps_done + t2.steps_done, model=model)
