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
        self.load(data)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self._iteration()
        return self

    def run(self, data):
        self.load(data)
        self.graph.replay()
        self.replays += 1
