"""DPF: the differentiable particle filter module with the reference's constructor, attributes, state_dict keys
and method signatures (reference DPFs.py:22-451).

`filtering_pos` is the hot path (DPFs.py:144-216).  Per timestep it launches a handful of fused sm_100a kernels
through libnfdpf (soft / OT resampling, coupling stacks, measurement + weight update) instead of the reference's
6-17 thousand ATen calls, keeps the per-step lists as Python lists stacked once at the end (the reference
re-concatenates them every step, O(T^2) copies), and never materialises a (B,N,N) or (P,C) tensor."""
import os

import numpy as np
import torch
import torch.nn as nn

from .losses import autoencoder_loss, pseudolikelihood_loss, pseudolikelihood_loss_nf, supervised_loss
from .model.models import (build_conditional_nf, build_decoder, build_decoder_cglow, build_encoder, build_encoder_cglow,
                           build_likelihood, build_particle_encoder, build_particle_encoder_cglow, build_transition_model,
                           measurement_model_cnf, measurement_model_cosine_distance, measurement_model_Gaussian,
                           measurement_model_NN, motion_update, nf_dynamic_model, proposal_likelihood, _FusedMeasurement)
from .resamplers.resamplers import resampler
from .utils import checkpoint_state, compute_normal_density, load_model, normalize_log_probs, particle_initialization

device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


class DPF(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.param = args
        self.NF, self.NFcond = args.NF_dyn, args.NF_cond
        self.measurement = args.measurement
        self.hidden_size = args.hiddensize
        self.state_dim = 2
        self.lr, self.alpha = args.lr, args.alpha
        self.seq_len, self.num_particle, self.batch_size = args.sequence_length, args.num_particles, args.batchsize
        self.labeledRatio = args.labeledRatio
        self.spring_force, self.drag_force = 0.1, 0.0075
        self.pos_noise, self.vel_noise = args.pos_noise, args.vel_noise
        self.NF_lr = args.NF_lr
        self.n_sequence = 2
        self.build_model()
        self.eps, self.scaling, self.threshold, self.max_iter = args.epsilon, args.scaling, args.threshold, args.max_iter
        self.resampler = resampler(self.param)
        # optional hooks: injected random draws (parity tests / benchmarks) and a gate override
        self.injected = None          # dict(init_particles, noise (B,T,N,2), offsets (B,T)) or None
        self.force_resample = None    # None = the reference's ESS gate; True / False = always / never
        self.rng_device = "cpu"       # "cpu": draws come from the CPU generator in the reference's order (seed-compatible);
                                      # "cuda": initial cloud, motion noise and resampling offsets are drawn in-kernel (Philox,
                                      #         no per-step H2D, CUDA-graph capturable) -- SURVEY 8(f2)
        self._fired_host, self._gates = [], None
        self.hoist_encoder = bool(getattr(args, "hoist_encoder", False))   # ONE encoder call on all B*T frames before
                                      # the loop instead of T calls on B frames (DPFs.py:177).  Identical in eval mode; in train
                                      # mode BatchNorm then normalises with B*T-frame batch statistics instead of per-step ones.
        self.dist_group = None        # a torch.distributed group: the ESS gate then uses the mean over EVERY shard's trajectories
                                      # (one 16-byte all-reduce per timestep) -- see distributed.py
        if getattr(args, "fast", False):   # --fast: host-free filter loop (device-side ESS gate + in-kernel random draws)
            self.rng_device = "cuda"

    # ------------------------------------------------------------------------------------------ construction
    def build_model(self):
        cglow = self.measurement == "CGLOW"
        self.encoder = (build_encoder_cglow if cglow else build_encoder)(self.hidden_size)
        self.decoder = (build_decoder_cglow if cglow else build_decoder)(self.hidden_size)
        self.build_particle_encoder = build_particle_encoder_cglow if cglow else build_particle_encoder
        self.particle_encoder = self.build_particle_encoder(self.hidden_size, self.state_dim)
        self.transition_model = build_transition_model(self.state_dim)  # built but unused, like the reference
        self.motion_update = motion_update
        self.nf_dyn = build_conditional_nf(self.n_sequence, 2 * self.state_dim, self.state_dim, init_var=0.01)
        self.cond_model = build_conditional_nf(self.n_sequence, 2 * self.state_dim + self.hidden_size, self.state_dim, init_var=0.01)
        if self.measurement == "CRNVP":
            self.cnf_measurement = build_conditional_nf(self.n_sequence, self.hidden_size, self.hidden_size, init_var=0.01, prior_std=2.5)
            self.measurement_model = measurement_model_cnf(self.particle_encoder, self.cnf_measurement)
        elif self.measurement == "cos":
            self.measurement_model = measurement_model_cosine_distance(self.particle_encoder)
        elif self.measurement == "NN":
            self.likelihood_est = build_likelihood(self.hidden_size, self.state_dim)
            self.measurement_model = measurement_model_NN(self.particle_encoder, self.likelihood_est)
        elif self.measurement == "gaussian":
            self.gaussian_distribution = torch.distributions.MultivariateNormal(torch.ones(self.hidden_size).to(device),
                                                                                100 * torch.eye(self.hidden_size).to(device))
            self.measurement_model = measurement_model_Gaussian(self.particle_encoder, self.gaussian_distribution)
        elif cglow:
            raise NotImplementedError("--measurement CGLOW is outside the accelerated hot path")
        self.prototype_density = compute_normal_density(pos_noise=self.pos_noise, vel_noise=self.vel_noise)
        self.optim = torch.optim.Adam(self.parameters(), lr=self.lr)
        self.optim_scheduler = torch.optim.lr_scheduler.MultiStepLR(self.optim, milestones=[30 * (1 + x) for x in range(10)], gamma=1.0)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, inputs, train=True):
        start_image, start_state, image, state, q, visible = inputs
        state, start_state = state.to(device), start_state.to(device)
        image = image.permute(0, 1, 4, 2, 3).to(device)
        if self.rng_device == "cuda":
            vel = state[:, :, 2:] + 4.0 * torch.randn(state[:, :, 2:].shape, device=state.device, dtype=state.dtype)
        else:
            vel = state[:, :, 2:] + torch.normal(0.0, 4.0, state[:, :, 2:].shape).to(device)
        (particle_list, particle_weight_list, noise_list, likelihood_list, init_weights_log, index_list, jac_list, prior_list,
         obs_likelihood) = self.filtering_pos(image, start_state, vel)
        mask = self.get_mask() if train else 1.0
        loss_sup, predictions = supervised_loss(particle_list, particle_weight_list, state, mask, train)
        loss_ae = autoencoder_loss(image, train, self.encoder, self.decoder)
        lamda1, lamda2, lamda3 = 1.0, 0.01, 2.0
        if self.param.trainType == "DPF":
            loss_pseud_lik = None
            total_loss = lamda1 * loss_sup + lamda3 * loss_ae
        elif self.param.trainType == "SDPF":
            if self.NF:
                loss_pseud_lik = pseudolikelihood_loss_nf(particle_weight_list, noise_list, likelihood_list, index_list, jac_list,
                                                          prior_list, self.param.block_length)
            else:
                loss_pseud_lik = pseudolikelihood_loss(particle_weight_list, noise_list, likelihood_list, index_list,
                                                       self.param.block_length, self.param.pos_noise, self.param.vel_noise)
            total_loss = lamda1 * loss_sup + lamda2 * loss_pseud_lik + lamda3 * loss_ae
        else:
            raise ValueError("Please select the training type in DPF (supervised learning) and SDPF (semi-supervised learning)")
        return (total_loss, loss_sup, loss_pseud_lik, loss_ae, predictions, particle_list, particle_weight_list, state, start_state,
                image, likelihood_list, noise_list, obs_likelihood)

    # ----------------------------------------------------------------------------------------------- hot path
    @property
    def fired(self):
        """Per-timestep ESS-gate decisions of the last filtering_pos call (list of bool).  With the device-side gate they are
        read back lazily -- asking for them is the only device-to-host synchronisation they ever cause."""
        if self._fired_host is None and self._gates is not None:
            self._fired_host = [bool(v) for v in self._gates.tolist()]
        return self._fired_host if self._fired_host is not None else []

    @fired.setter
    def fired(self, value):
        self._fired_host, self._gates = list(value), None

    def _rng_state(self, dev):
        """Device int64[2] = {seed, step counter} of the in-kernel Philox draws (rng_device = "cuda").  Eager calls re-seed it from
        torch's CPU generator (so `torch.manual_seed(s)` makes a run reproducible, like the reference's CPU draws) and restart the
        counter; while a CUDA graph is being captured nothing is written from the host: replays keep advancing the device-side
        counter, so every replay sees fresh noise."""
        st = getattr(self, "_rng", None)
        if st is None or st.device != dev:
            st = torch.zeros(2, dtype=torch.int64, device=dev)
            object.__setattr__(self, "_rng", st)     # plain attribute: not a buffer, so state_dict keys stay the reference's
        if not torch.cuda.is_current_stream_capturing():
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            st.copy_(torch.tensor([seed, 0], dtype=torch.int64), non_blocking=True)
        return st

    def filtering_pos(self, obs, start_state_vs, vel_input):
        """The per-timestep particle update (reference DPFs.py:144-216).  `obs` is (B,T,3,H,W) images, or (B,T,h)
        precomputed encodings when `self.encoder` is an Identity (benchmarks exclude the CNN).

        Host-free where the reference is not: the whole-batch ESS gate (DPFs.py:163-165) is a device flag the resampling kernels
        read (no sync per step) unless the reference's CPU-generator draw order has to be reproduced (rng_device = "cpu" with
        nothing injected: the offsets are only drawn when the gate fires, so the host must know); the per-step outputs are
        written by the kernels straight into pre-allocated (T,B,...) buffers and returned as transposed views (the reference
        re-concatenates every list every step); the supervised loss' per-step predictions (losses.py:22) come from a fused
        kernel and ride along on the returned particle list (see losses.supervised_loss)."""
        from . import ops
        start_state, vel = start_state_vs[:, :2], start_state_vs[:, 2:]
        B, N, T, inj = start_state.shape[0], self.num_particle, self.seq_len, self.injected or {}
        dev = start_state.device
        device_rng = self.rng_device == "cuda"
        rng = self._rng_state(dev) if device_rng else None
        if "init_particles" in inj:   # injected cloud: no host RNG, no H2D copy (graph-capturable)
            particles = inj["init_particles"]
            init_weights_log = torch.log(torch.ones(B, N, device=dev) / N)
        elif device_rng:
            particles = ops.init_particles(start_state, self.param.width, N, rng, self.param.init_with_true_state)
            init_weights_log = torch.log(torch.ones(B, N, device=dev) / N)
        else:
            particles, init_weights_log = particle_initialization(start_state, self.param.width, N, self.state_dim,
                                                                  init_with_true_state=self.param.init_with_true_state)
        _, particle_probs, _, ess_inv = _weight_norm(init_weights_log)
        fused = isinstance(self.measurement_model, _FusedMeasurement)
        soft = self.param.resampler_type == "soft"
        f32 = dict(dtype=torch.float32, device=dev)
        buf = {"particles": torch.empty(T, B, N, 2, **f32), "probs": torch.empty(T, B, N, **f32), "lki": torch.empty(T, B, N, **f32),
               "pred": torch.empty(T, B, 2, **f32)}
        if "noise" not in inj:
            buf["noise"] = torch.empty(T, B, N, 2, **f32)
        if soft:
            buf["index"] = torch.empty(T, B, N, dtype=torch.int64, device=dev)
        if self.NF:
            buf["jac"], buf["prior"] = torch.empty(T, B, N, **f32), torch.empty(T, B, N, **f32)
        steps = {k: [] for k in ("particles", "probs", "lki", "pred", "jac", "prior")}
        row_sums = []
        # gate mode: constant (force_resample), host rule (reference RNG order, one sync per step) or device rule (no sync)
        host_rule = self.force_resample is None and soft and not device_rng and "offsets" not in inj
        device_rule = self.force_resample is None and not host_rule
        self._gates = torch.empty(T, dtype=torch.int32, device=dev) if device_rule else None
        self._fired_host = None if device_rule else []
        identity_idx = None
        enc_all = None
        if self.hoist_encoder and not isinstance(self.encoder, nn.Identity):
            enc_all = self.encoder(obs.reshape((B * T,) + tuple(obs.shape[2:])).float()).reshape(B, T, -1)

        # per-step slices of (B,T,...) inputs are strided views, and every kernel call would first copy them: transpose once
        vel_steps = vel_input.transpose(0, 1).contiguous()
        obs_steps = obs.transpose(0, 1).contiguous() if (enc_all is None and obs.dim() == 3) else None   # precomputed encodings
        enc_steps = enc_all.transpose(0, 1).contiguous() if enc_all is not None else None

        def put(key, step, t):   # data the kernels could not write in place (non-fused paths) is copied into the list buffer
            if t.data_ptr() != buf[key][step].data_ptr():
                buf[key][step].copy_(t.detach())

        for step in range(T):
            gate = None
            if self.dist_group is not None and self.force_resample is None:
                from .distributed import global_ess_mean
                ess_inv = global_ess_mean(ess_inv, self.dist_group)     # (1,): the whole-batch mean over all shards
            if device_rule:
                want_off = soft and device_rng and "offsets" not in inj
                if ess_inv.numel() == B:
                    gate, off_dev = ops.ess_gate(ess_inv, B, N, None, rng, device_rng, want_off, self._gates[step:step + 1])
                else:   # global mean of a sharded run: the rule on one value, then the per-trajectory offsets / counter advance
                    gate, _ = ops.ess_gate(ess_inv, 1, N, None, None, False, False, self._gates[step:step + 1])
                    _, off_dev = ops.ess_gate(None, B, N, True, rng, device_rng, want_off) if device_rng else (None, None)
                fire = True        # the kernels decide
            else:
                off_dev = None
                if self.force_resample is None:
                    fire = bool(ess_inv.mean() < 0.5 * N)          # whole-batch ESS gate, DPFs.py:163-165 (one D2H sync)
                else:
                    fire = bool(self.force_resample)
                    if device_rng:                                  # keep the step counter of the device RNG moving
                        _, off_dev = ops.ess_gate(None, B, N, fire, rng, True, soft and fire and "offsets" not in inj)
                self._fired_host.append(fire)
            if fire and soft:   # one kernel: scan + search + gather + renormalise (+ log of the new weights, DPFs.py:167)
                off = inj["offsets"][:, step] if "offsets" in inj else off_dev
                particles, probs_res, index_p, logw_prev = self.resampler.resampling(
                    particles, particle_probs, random_offset=off, want_log=True, gate=gate, out={"index": buf["index"][step]},
                    **self.resampler.kargs)
            elif fire:
                k = self.resampler.kargs
                particles = ops.ot_resample(particles, particle_probs.log(), k["eps"], k["scaling"], k["threshold"], k["max_iter"], gate)
                if gate is not None:
                    probs_res, logw_prev = ops.gate_weights(particle_probs, gate)
                else:
                    probs_res = torch.full_like(particle_probs, 1.0 / N)
                    logw_prev = probs_res.log()
            else:
                logw_prev = particle_probs.log()
                if soft:
                    if identity_idx is None:
                        identity_idx = torch.arange(B * N, device=dev, dtype=torch.int64).reshape(B, N)
                    buf["index"][step].copy_(identity_idx)
            noise = inj["noise"][:, step] if "noise" in inj else None
            pred = None
            if enc_steps is not None:
                encodings = enc_steps[step]
            else:
                encodings = self.encoder((obs_steps[step] if obs_steps is not None else obs[:, step]).float())
            if fused:
                # ---- fused step: 6 libnfdpf launches (motion+moments, 3 coupling stacks, densities, measurement+update)
                if noise is None and not device_rng:
                    noise = torch.normal(mean=0.0, std=self.pos_noise, size=(B, N, 2)).to(dev, non_blocking=True)
                    put("noise", step, noise)
                ctx_phys = torch.empty(B, 4, **f32) if self.NF else None
                plain = not self.NF and not self.NFcond             # the moved cloud itself is the step's particle set
                o = {"noise": buf["noise"][step]} if noise is None else {}
                if plain:
                    o["moved"] = buf["particles"][step]
                particles_physical, noise = ops.motion_moments(particles, vel, noise, ctx_phys, 0, rng, self.pos_noise, o)
                vel = vel_steps[step]
                if self.NF:
                    o = {"log_det": buf["jac"][step]}
                    if not self.NFcond:
                        o["y"] = buf["particles"][step]
                    particles_dynamical, jac = self.nf_dyn.run_stack(particles_physical, row_ctx=ctx_phys, inverse=True, neg_logdet=True, out=o)
                else:
                    particles_dynamical, jac = particles_physical, None
                mo = {"lki": buf["lki"][step], "probs": buf["probs"][step], "pred": buf["pred"][step]}
                if self.NFcond:
                    ctx_prop = torch.empty(B, self.hidden_size + 4, **f32)
                    # proposal sees a detached encoding, models.py:360-361: copied into the context row by the moments launch
                    ops.row_moments(particles_dynamical, ctx_prop, self.hidden_size, head=encodings.detach())
                    propose_particle, jac_prop = self.cond_model.run_stack(particles_dynamical, row_ctx=ctx_prop, inverse=True, neg_logdet=True,
                                                                           out={"y": buf["particles"][step]})
                    # the proposal has three consumers (dynamics flow, measurement + prediction, next step): one alias each, their
                    # gradients meet in one summing launch instead of chained autograd adds
                    x_back, x_meas, propose_particle = ops.fanout(propose_particle, 3)
                    if self.NF:   # push the proposal back through the dynamics flow (context: moments of the physical cloud)
                        back, jac_back = self.nf_dyn.run_stack(x_back, row_ctx=ctx_phys, inverse=False, neg_logdet=True)
                    else:
                        back, jac_back = x_back, None
                    prior_log, propose_log = ops.proposal_terms(back, particles_physical, noise, jac_back, jac, jac_prop, self.pos_noise,
                                                                {"prior": buf["prior"][step]} if self.NF else None)
                    lki_log, logw, particle_probs, row_sum, ess_inv, pred = self.measurement_model.forward_update(
                        encodings, x_meas, logw_prev, prior_log, propose_log, out=mo, want_pred=True)
                else:             # prior == proposal density: the two cancel exactly in DPFs.py:187
                    propose_particle = particles_dynamical
                    if self.NF:
                        _, prior_log = ops.proposal_terms(particles_physical, particles_physical, noise, None, jac, None, self.pos_noise,
                                                          {"propose": buf["prior"][step]})
                    lki_log, logw, particle_probs, row_sum, ess_inv, pred = self.measurement_model.forward_update(
                        encodings, propose_particle, logw_prev, None, None, out=mo, want_pred=True)
            else:
                particles_physical, noise = self.motion_update(particles, vel, pos_noise=self.pos_noise, noise=noise)
                if "noise" in buf:
                    put("noise", step, noise)
                vel = vel_input[:, step, :]
                particles_dynamical, jac = nf_dynamic_model(self.nf_dyn, particles_physical, particle_probs.shape, NF=self.NF)
                propose_particle, lki_log, prior_log, propose_log = proposal_likelihood(
                    self.cond_model, self.nf_dyn, self.measurement_model, particles_dynamical, particles_physical, encodings, noise, jac,
                    self.NF, self.NFcond, prototype_density=self.prototype_density)
                logw, particle_probs, row_sum, ess_inv = _weight_update(logw_prev, lki_log, prior_log, propose_log,
                                                                        out={"probs": buf["probs"][step]})
                put("lki", step, lki_log)
                if self.NF:
                    put("jac", step, jac)
                    put("prior", step, prior_log)
            particles = propose_particle
            put("particles", step, particles)
            row_sums.append(row_sum)
            if pred is None:     # non-fused paths: the prediction of losses.py:22 as its own kernel (fused paths: measurement epilogue)
                pred = ops.weighted_mean(particles, particle_probs, {"pred": buf["pred"][step]})
            for k, v in (("particles", particles), ("probs", particle_probs), ("lki", lki_log), ("pred", pred)):
                steps[k].append(v)
            if self.NF:
                steps["jac"].append(jac)
                steps["prior"].append(prior_log)
        obs_likelihood = torch.stack(row_sums).sum() / (B * N)            # sum_t mean_{b,n} logw, DPFs.py:191
        view = lambda k: ops.list_view(buf[k], steps[k])
        particle_list, weight_list = view("particles"), view("probs")
        # the fused per-step predictions ride along: losses.supervised_loss(particle_list, weight_list, ...) picks them up instead
        # of multiplying and reducing the (B,T,N,2) lists again
        particle_list._nfdpf_pred = (view("pred"), weight_list)
        noise_list = inj["noise"] if "noise" in inj else buf["noise"].transpose(0, 1)
        if soft:
            index_list = buf["index"].transpose(0, 1)
        else:   # OT resampling keeps every particle in place: identity ancestors at every step (resamplers.py:68-70), as a view
            index_list = torch.arange(B * N, device=dev, dtype=torch.int64).reshape(B, 1, N).expand(B, T, N)
        return (particle_list, weight_list, noise_list, view("lki"), init_weights_log, index_list,
                view("jac") if self.NF else None, view("prior") if self.NF else None, obs_likelihood)

    def get_mask(self):
        n1 = int(self.batch_size * self.seq_len * self.labeledRatio)
        arr = np.array([0] * (self.batch_size * self.seq_len - n1) + [1] * n1)
        np.random.shuffle(arr)
        return torch.tensor(arr.reshape(self.batch_size, self.seq_len)).to(device)

    # ------------------------------------------------------------------------- trainer glue (out of hot path)
    def _run_epoch(self, loader, train):
        sup, ae, last = [], [], None
        for inputs in loader:
            out = self.forward(inputs, train=train)
            if train:
                self.zero_grad()
                out[0].backward()
                self.optim.step()
            sup.append(out[1].detach().cpu().numpy())
            ae.append(out[3].detach().cpu().numpy())
            last = out
        return sup, ae, last

    def pretrain_ae(self, train_loader, valid_loader, start_epoch=-1, epoch_num=100, logger=None):
        best, ckpt_ae = 1e10, None
        for epoch in range(start_epoch + 1, epoch_num):
            for phase, loader in (("train", train_loader), ("val", valid_loader)):
                self.train(phase == "train")
                losses = []
                with torch.set_grad_enabled(phase == "train"):
                    for inputs in loader:
                        img = inputs[2].permute(0, 1, 4, 2, 3).reshape(-1, 3, 128, 128).to(device)
                        loss = nn.functional.mse_loss(self.decoder(self.encoder(img)), img)
                        if phase == "train":
                            self.zero_grad()
                            loss.backward()
                            self.optim.step()
                        losses.append(loss.detach().cpu().numpy())
                print(f"{phase} AE: Epoch: {epoch}, loss: {np.mean(losses)}")
            if logger is not None:
                logger.add_scalar("PretrainAE_loss_eval/loss", np.mean(losses), epoch)
            if np.mean(losses) < best:
                best = np.mean(losses)
                ckpt_ae = {"model": self.state_dict(), "optim": self.optim.state_dict()}
                torch.save(ckpt_ae, "./model/ae_pretrain.pth")
        if ckpt_ae is not None:
            self.load_state_dict(ckpt_ae["model"])
            self.optim.load_state_dict(ckpt_ae["optim"])

    def e2e_train(self, train_loader, valid_loader, start_epoch=-1, epoch_num=100, logger=None, run_id=None):
        best, history = 1e10, []
        if self.param.load_pretrainModel:
            self.load_state_dict(torch.load("./model/ae_pretrain.pth")["model"])
        for epoch in range(start_epoch + 1, epoch_num):
            self.train()
            sup, ae, last = self._run_epoch(train_loader, True)
            self.optim_scheduler.step()
            if logger is not None:
                logger.add_scalar("Sup_loss/loss", np.mean(sup), epoch)
            print(f"End-to-end loss: epoch: {epoch}, loss: {np.mean(sup)}, loss_ae: {np.mean(ae)}, obs_likelihood: {last[-1]}")
            self.eval()
            with torch.no_grad():
                sup_eval, _, last = self._run_epoch(valid_loader, False)
            eval_mean = np.mean(sup_eval)
            if logger is not None:
                logger.add_scalar("Sup_loss_eval/loss", eval_mean, epoch)
            print(f"End-to-end loss evaluation: epoch: {epoch}, loss: {eval_mean}, obs_likelihood: {last[-1]}", self.NF)
            history.append(eval_mean)
            np.save(os.path.join("logs", run_id, "data", "eval_loss_epoch.npy"), history)
            if eval_mean < best:
                best = eval_mean
                _dump(os.path.join("logs", run_id, "data", "eval_result_best.npz"), last, loss=sup_eval)
                torch.save(checkpoint_state(self, epoch), os.path.join("logs", run_id, "models", "e2e_model_bestval_e2e.pth"))

    def load_model(self, file_name):
        ckpt = torch.load(file_name)
        load_model(self, ckpt)
        print(f"Load epcoh: {ckpt['epoch']}")

    def train_val(self, train_loader, valid_loader, run_id):
        from torch.utils.tensorboard import SummaryWriter
        for d in ("result", "model", "checkpoint", "logger"):
            os.makedirs(d, exist_ok=True)
        logger = SummaryWriter("./logger")
        if self.param.resume:
            self.load_model("./model/e2e_model_bestval_e2e.pth")
        if self.param.pretrain_ae:
            self.pretrain_ae(train_loader, valid_loader, start_epoch=-1, epoch_num=300, logger=logger)
        if self.param.e2e_train:
            self.e2e_train(train_loader, valid_loader, start_epoch=-1, epoch_num=self.param.num_epochs, logger=logger, run_id=run_id)

    def testing(self, test_loader, run_id, model_path="./model/e2e_model_bestval_e2e.pth"):
        if self.param.testing:
            self.load_model(os.path.join(model_path, "e2e_model_bestval_e2e.pth"))
        self.eval()
        with torch.no_grad():
            sup_eval, _, last = self._run_epoch(test_loader, False)
        np.save(os.path.join("logs", run_id, "data", "test_loss_epoch.npy"), sup_eval)
        print(f"End-to-end loss testing: loss: {np.mean(sup_eval)}")
        _dump(os.path.join("logs", run_id, "data", "test_result.npz"), last, images=last[9].detach().cpu().numpy(),
              noise=last[11].detach().cpu().numpy())


# ------------------------------------------------------------------------------------------------- helpers
def _weight_norm(logw):
    from . import ops
    return ops.weight_update(logw)


def _weight_update(logw_prev, lki, prior, propose, out=None):
    from . import ops
    return ops.weight_update(logw_prev, lki, prior, propose, 1e-12, out)


def _dump(path, out, **extra):
    np.savez(path, particle_list=out[5].detach().cpu().numpy(), particle_weight_list=out[6].detach().cpu().numpy(),
             likelihood_list=out[10].detach().cpu().numpy(), pred=out[4].detach().cpu().numpy(), state=out[7].detach().cpu().numpy(),
             **extra)
