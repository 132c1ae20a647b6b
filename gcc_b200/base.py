"""Plumbing shared by the four model wrappers (Pix2PixModel, MobileCycleGANModel, SRGAN, SAGANModel): everything a
CUDA-graph replay, a data-parallel launch or a resume needs to find WITHOUT knowing which model it drives.

  * device-resident step scalars: the learning rate of every optimizer arena and the teacher-EMA factor
    ``opt.ema_beta`` (models/Pix2Pix.py:503-508, 602-611) live in device memory; ``refresh_device_scalars()`` copies
    the current host values there.  A captured iteration reads the device copies, so `update_learning_rate(epoch)`
    and a changed ``opt.ema_beta`` reach a replayed graph exactly as they reach the eager call sequence.
  * data parallel: ``broadcast_parameters()`` replicates rank 0's parameters / BatchNorm buffers (the reference's
    train.py sets no seed, so every rank would otherwise draw its own initial weights) and de-correlates the
    dropout streams of the ranks.
  * resume: ``resume_state()`` / ``load_resume_state()`` hold what the reference's checkpoints (G / D state dicts
    only, models/Pix2Pix.py:636-658) leave out: Adam moments and step counters, the un-registered transform convs,
    the EMA scalars, dropout seeds, scheduler positions and the teacher.
"""
import torch


def capturing():
    """True while the current CUDA stream records a graph (host-side scalar refreshes must not be captured)."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def dist_on():
    d = torch.distributed
    return d.is_available() and d.is_initialized() and d.get_world_size() > 1


class GccModelMixin:
    # ---- to be provided by the model class -------------------------------------------------------------------
    def _gcc_arenas(self):
        """{name: ParamArena} of every arena the model owns."""
        raise NotImplementedError

    def _gcc_optimizers(self):
        """{name: optimizer facade} (zero_grad / step / param_groups)."""
        raise NotImplementedError

    def _gcc_nets(self):
        """{name: net} of the nn.Modules holding BatchNorm buffers / dropout seeds."""
        raise NotImplementedError

    # ---- device scalars ---------------------------------------------------------------------------------------
    def _base_init(self):
        self._ema_states = {}
        self._ema_beta_host = float(self.opt.ema_beta)
        self._ema_beta_dev = torch.full((), self._ema_beta_host, dtype=torch.float32, device=self.device)
        if getattr(self.opt, "sync_bn", False) and dist_on():
            from . import ops
            ops.GLOBAL_BATCH_SYNC = True

    def refresh_device_scalars(self):
        """Host -> device copies of the values that change between iterations outside the step itself.  Must run
        outside a stream capture (GraphedIteration.run() calls it before every replay)."""
        b = float(self.opt.ema_beta)
        if b != self._ema_beta_host:
            self._ema_beta_host = b
            self._ema_beta_dev.fill_(b)
        for o in self._gcc_optimizers().values():
            lr = o.param_groups[0]["lr"]
            for a in (o.arenas_of() if hasattr(o, "arenas_of") else [o.arena]):
                a.set_lr(lr)

    def _ema_update(self, key, diff):
        """Teacher side of get_D_arch_diff (models/Pix2Pix.py:503-508): state <- beta * diff + (1 - beta) * state, in
        ONE persistent device scalar per key so that a captured CUDA graph carries it across replays; beta is read
        from device memory."""
        if not capturing():
            b = float(self.opt.ema_beta)
            if b != self._ema_beta_host:
                self._ema_beta_host = b
                self._ema_beta_dev.fill_(b)
        st = self._ema_states.get(key)
        d = diff.detach()
        if st is None:
            st = self._ema_states[key] = d.clone()
        else:
            bt = self._ema_beta_dev
            st.copy_(bt * d + (1.0 - bt) * st)
        return st

    # ---- norm layers / counters -------------------------------------------------------------------------------
    def _gcc_norm_layers(self):
        out = []
        for net in self._gcc_nets().values():
            out += list(getattr(net, "_norms", []))
        return out

    # ---- data parallel ----------------------------------------------------------------------------------------
    def broadcast_parameters(self, src=0):
        """Replicate rank `src`'s parameters and BatchNorm buffers; rank-dependent dropout streams."""
        if not dist_on():
            return
        dist = torch.distributed
        for a in self._gcc_arenas().values():
            dist.broadcast(a.P, src)
            a.mark_dirty()
        for l in self._gcc_norm_layers():
            if l.running_mean is not None:
                dist.broadcast(l.running_mean, src)
                dist.broadcast(l.running_var, src)
        rank = dist.get_rank()
        for net in self._gcc_nets().values():
            seed = getattr(net, "seed", None)
            if torch.is_tensor(seed) and not getattr(net, "_seed_ranked", False):
                seed.add_(rank * 0x632BE59BD9B4E019 & 0x3FFFFFFFFFFFFFFF)
                net._seed_ranked = True

    # ---- resume -----------------------------------------------------------------------------------------------
    def finish_pending_steps(self):
        """Complete deferred data-parallel optimizer steps before parameters are read from outside the step."""
        for o in self._gcc_optimizers().values():
            if hasattr(o, "finish"):
                o.finish()

    def resume_state(self, with_teacher=True):
        self.finish_pending_steps()
        cpu = lambda t: t.detach().to("cpu", copy=True)
        st = {"arenas": {}, "buffers": {}, "ema": {k: cpu(v) for k, v in self._ema_states.items()},
              "ema_beta": float(self.opt.ema_beta), "seeds": {}, "lr": {}, "sched": []}
        for name, a in self._gcc_arenas().items():
            st["arenas"][name] = {"P": cpu(a.P), "M": cpu(a.M), "V": cpu(a.V), "hyper": cpu(a.hyper),
                                  "specs": [(n, tuple(s), k) for n, s, k in a.specs]}
        for name, net in self._gcc_nets().items():
            if hasattr(net, "sync_counters"):
                net.sync_counters()
            st["buffers"][name] = {k: cpu(v) for k, v in net.named_buffers()}
            seed = getattr(net, "seed", None)
            if torch.is_tensor(seed):
                st["seeds"][name] = cpu(seed)
        for name, o in self._gcc_optimizers().items():
            st["lr"][name] = dict(o.param_groups[0])
        st["sched"] = [getattr(s, "epoch", 0) for s in getattr(self, "schedulers", [])]
        extra = getattr(self, "_gcc_extra_state", None)
        if extra is not None:
            st["extra"] = extra()
        if with_teacher and getattr(self, "teacher_model", None) is not None:
            st["teacher"] = self.teacher_model.resume_state(with_teacher=False)
        return st

    def load_resume_state(self, st):
        with torch.no_grad():
            for name, a in self._gcc_arenas().items():
                s = st["arenas"][name]
                if [(n, tuple(sh), k) for n, sh, k in a.specs] != [(n, tuple(sh), k) for n, sh, k in s["specs"]]:
                    raise RuntimeError("resume state of arena %r does not match this model's layout" % name)
                for key in ("P", "M", "V", "hyper"):
                    getattr(a, key).copy_(s[key].to(a.device))
                a.lr = float(s["hyper"][0])
                a.mark_dirty()
            for name, net in self._gcc_nets().items():
                own = dict(net.named_buffers())
                for k, v in st["buffers"].get(name, {}).items():
                    own[k].copy_(v.to(own[k].device))
                for l in getattr(net, "_norms", []):
                    if l.mode == "bn":
                        l.num_batches = int(l.nbt)
                if name in st["seeds"]:
                    net.seed.copy_(st["seeds"][name].to(net.seed.device))
            self._ema_states = {k: v.to(self.device).clone() for k, v in st["ema"].items()}
        self.opt.ema_beta = st["ema_beta"]
        for name, o in self._gcc_optimizers().items():
            o.param_groups[0].update(st["lr"][name])
        for s, e in zip(getattr(self, "schedulers", []), st["sched"]):
            s.epoch = e
        if "extra" in st and hasattr(self, "_gcc_load_extra_state"):
            self._gcc_load_extra_state(st["extra"])
        self.refresh_device_scalars()
        if "teacher" in st and getattr(self, "teacher_model", None) is not None:
            self.teacher_model.load_resume_state(st["teacher"])

    # ---- checkpoint hooks -------------------------------------------------------------------------------------
    RESUME_KEY = "gcc_b200"

    def _ckpt_add_resume(self, ckpt):
        """The reference's checkpoint dict plus ONE extra entry (ignored by the reference's load_models, which reads
        its own keys only): everything needed to continue training exactly where it stopped."""
        ckpt[self.RESUME_KEY] = self.resume_state()
        return ckpt

    def _ckpt_load_resume(self, ckpt, load_discriminator, resume=None):
        want = (self.RESUME_KEY in ckpt and load_discriminator) if resume is None else resume
        if want:
            if self.RESUME_KEY not in ckpt:
                raise RuntimeError("checkpoint has no %r entry (written by the reference?): cannot resume" % self.RESUME_KEY)
            st = ckpt[self.RESUME_KEY]
            if getattr(self, "teacher_model", None) is None:
                st = {k: v for k, v in st.items() if k != "teacher"}
            self.load_resume_state(st)
