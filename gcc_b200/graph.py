"""CUDA-graph capture of one whole GCC iteration (teacher step + student step + gate step) of any of the four
model wrappers (train.py:144-151):

    model.set_input(train); model.optimize_parameters()
    model.set_input(val);   model.clipping_mask_alpha(); model.optimizer_netD_arch()      # gated discriminator only

The iteration launches ~1000 small-to-large kernels from Python; captured once and replayed, the host cost per
iteration drops to the input copies and one graph launch.  Everything the step needs that changes between
iterations lives in device memory: Adam step counters, dropout seeds, the teacher EMA scalars, loss scalars, the
CycleGAN image pools (`cyclegan.DeviceImagePool`) -- and the learning rates and ``opt.ema_beta``, whose host values
``run()`` copies to the device before every replay (`GccModelMixin.refresh_device_scalars`), so a replay is the same
program as the eager call sequence, including after ``update_learning_rate(epoch)``.
tests/test_graph_replay_gpu.py checks replay == eager for all of this.
"""
import torch

from . import base


def _is_batch_tensor(v):
    return torch.is_tensor(v) and v.is_floating_point()


class GraphedIteration:
    def __init__(self, model):
        self.model = model
        self.static_train = self.static_val = None
        self.graph = None
        self.segments = []
        self.replays = 0
        self.has_arch = bool(getattr(model.opt, "darts_discriminator", False)) and model.teacher_model is not None
        self._bn_delta = []
        self.pace = True
        self._async = {}
        self._pace_ev = torch.cuda.Event()

    # ------------------------------------------------------------------ inputs
    def _make_static(self, batch):
        dev = self.model.device
        return {k: (torch.empty(v.shape, dtype=v.dtype, device=dev) if _is_batch_tensor(v) else v) for k, v in batch.items()}

    @staticmethod
    def _fill(static, batch):
        for k, t in static.items():
            if _is_batch_tensor(t):
                t.copy_(batch[k], non_blocking=True)

    def load(self, train, val=None):
        self._fill(self.static_train, train)
        if self.has_arch:
            self._fill(self.static_val, val)

    def _iteration(self):
        m = self.model
        m.set_input(self.static_train)
        m.optimize_parameters()
        if self.has_arch:
            m.set_input(self.static_val)
            m.clipping_mask_alpha()
            m.optimizer_netD_arch()
        for mm in self._models():          # an iteration ends with no optimizer step in flight
            mm.finish_pending_steps()

    def _models(self):
        return [m for m in (self.model, self.model.teacher_model) if m is not None]

    def _refresh(self):
        for m in self._models():
            m.refresh_device_scalars()

    # ------------------------------------------------------------------ capture
    def capture(self, train, val=None, warmup=2):
        from . import pix2pix
        if self.has_arch and val is None:
            raise ValueError("a gated-discriminator iteration needs a validation batch for the gate step")
        for m in self._models():
            if hasattr(m, "use_device_pools"):
                m.use_device_pools()          # host-random ImagePool -> device-resident pool (capturable)
        self.static_train = self._make_static(train)
        self.static_val = self._make_static(val) if self.has_arch else None
        self.load(train, val)
        self._refresh()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):      # also creates every lazily allocated persistent scalar
                self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        layers = [l for m in self._models() for l in m._gcc_norm_layers()]
        before = [l.num_batches for l in layers]
        if not base.dist_on():
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._iteration()
            self.segments = [(self.graph, None)]
        else:
            # Data parallel: a chain of graphs cut at every gradient exchange; run() replays a segment, launches the NCCL
            # all-reduce of that optimizer group's flat gradient arena eagerly on the same stream, replays the next
            # segment (which starts with the Adam step).  All segments share one memory pool and are always replayed
            # in capture order.  (Capturing the collectives as graph nodes was tried twice in round 2 -- global and
            # thread-local capture mode -- and dead-locks against ProcessGroupNCCL's watchdog thread with torch 2.11 /
            # NCCL 2.28: not offered.)
            pool = torch.cuda.graph_pool_handle()
            state = {"g": None}

            def begin():
                state["g"] = torch.cuda.CUDAGraph()
                state["g"].capture_begin(pool=pool)

            def cut(arena, action="allreduce"):
                state["g"].capture_end()
                self.segments.append((state["g"], (arena, action)))
                begin()

            with torch.cuda.stream(side):
                begin()
                pix2pix._graph_segmenter = cut
                try:
                    self._iteration()
                finally:
                    pix2pix._graph_segmenter = None
                state["g"].capture_end()
                self.segments.append((state["g"], None))
            self.graph = self.segments[0][0]
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # train-mode BatchNorm counts its forwards on the host (num_batches_tracked): the capture pass ran the Python
        # code without executing a kernel (undo its count), every replay adds the captured count
        self._bn_delta = [(l, l.num_batches - b) for l, b in zip(layers, before) if l.num_batches != b]
        for l, b in zip(layers, before):
            l.num_batches = b
        return self

    # ------------------------------------------------------------------ replay
    def run(self, train, val=None):
        from . import pix2pix
        self.load(train, val)
        self._refresh()
        for g, boundary in self.segments:
            g.replay()
            if boundary is not None:
                arena, action = boundary
                if action == "allreduce":
                    pix2pix._allreduce_grads(arena)
                elif action == "allreduce_async":        # deferred step: the exchange overlaps the next segments
                    self._async[id(arena)] = pix2pix._allreduce_grads(arena, asynchronous=True)
                else:                                    # "wait": the next segment starts with that arena's Adam step
                    self._async.pop(id(arena)).wait()
        for l, d in self._bn_delta:
            l.num_batches += d
        self.replays += 1
        if len(self.segments) > 1 and self.pace:
            # data parallel with eager collectives between graph segments: keep the host at most one iteration ahead
            # (measured on 4 and 8 GPUs in round 1: an unpaced host, with all iterations' segments and collectives
            # queued at once, runs 1-4 ms per iteration slower)
            self._pace_ev.record()
            self._pace_ev.synchronize()
