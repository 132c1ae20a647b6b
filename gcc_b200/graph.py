"""CUDA-graph capture of one whole GCC iteration (teacher step + student step + gate step).

The iteration launches ~1000 small-to-large kernels from Python; captured once and replayed, the host cost
per iteration drops to four input copies and one graph launch.  Everything the step needs that changes between
iterations lives in device memory (Adam step counters and learning rates, dropout seeds, the teacher EMA scalar,
loss scalars), so a replay is numerically the same program as the eager call sequence of train.py:144-151.
"""
import torch


class GraphedIteration:
    def __init__(self, model, batch, size=256, device=None):
        self.model = model
        dev = device or model.device
        self.static = {k: torch.zeros(batch, 3, size, size, dtype=torch.float32, device=dev)
                       for k in ("A", "B", "vA", "vB")}
        self.graph = None
        self.replays = 0

    def _iteration(self):
        m, s = self.model, self.static
        m.set_input({"A": s["A"], "B": s["B"], "A_paths": "", "B_paths": ""})
        m.optimize_parameters()
        m.set_input({"A": s["vA"], "B": s["vB"], "A_paths": "", "B_paths": ""})
        m.clipping_mask_alpha()
        m.optimizer_netD_arch()

    def load(self, data):
        """data: dict with 'A','B' (train batch) and 'vA','vB' (validation batch for the gate step)."""
        for k, t in self.static.items():
            t.copy_(data[k], non_blocking=True)

    def capture(self, data, warmup=3):
        from . import pix2pix
        self.load(data)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if not pix2pix._dist_on():
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._iteration()
            self.segments = [(self.graph, None)]
            return self
        # Data parallel: the iteration is captured as a chain of graphs cut at every gradient exchange (five per
        # iteration); run() replays a segment, launches the NCCL all-reduce of that optimizer group's flat
        # gradient arena eagerly on the same stream, replays the next segment (which starts with the Adam step).
        # All segments share one memory pool and are always replayed in capture order.
        self.segments = []
        pool = torch.cuda.graph_pool_handle()
        state = {"g": None}

        def begin():
            state["g"] = torch.cuda.CUDAGraph()
            state["g"].capture_begin(pool=pool)

        def cut(arena):
            state["g"].capture_end()
            self.segments.append((state["g"], arena))
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
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = self.segments[0][0]
        return self

    def run(self, data):
        from . import pix2pix
        self.load(data)
        for g, arena in self.segments:
            g.replay()
            if arena is not None:
                pix2pix._allreduce_grads(arena)
        self.replays += 1
