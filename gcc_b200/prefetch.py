"""Input pipeline step in front of ``set_input`` (data/aligned_dataset.py:27-56 -> models/Pix2Pix.py:453-458): keeps
the 1000+ images/s step fed from host memory.

``Prefetcher(batches, device)`` wraps any iterable of batch dicts (what the reference's DataLoader yields: float32
NCHW tensors in [-1, 1] plus path strings) -- or of tuples of such dicts, e.g. (train batch, validation batch) of one
GCC iteration -- and yields the same structure with the tensors already on the device:

  * a background thread stages each tensor in a PINNED buffer (tensors that already are pinned are copied from where
    they lie) and issues the host->device copy of batch i+1 on a dedicated copy stream while the step of batch i
    computes (``depth`` slots on both sides, events in both directions: the copy stream does not overwrite a device
    slot before the consumer's stream is done with it, and the consumer's stream waits for the copy);
  * ``dtype=torch.bfloat16`` converts float32 tensors to bf16 while staging (the network rounds its input to bf16
    anyway: same values, half the PCIe bytes).  Off by default: measured on the pix2pix step (100 MB of fp32 input per
    iteration), the host-side conversion costs ~25 ms per iteration on one core -- more than the overlapped PCIe copy
    it saves -- so it only pays when the host has cores to spare and the link is the bottleneck.

The device tensors are NCHW like the host ones; ``set_input`` / ``GraphedIteration.run`` accept them unchanged.
"""
import queue
import threading

import torch


class Prefetcher:
    def __init__(self, batches, device, depth=2, dtype=None):
        self.src = iter(batches)
        self.device = torch.device(device)
        self.depth = depth
        self.dtype = dtype
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * depth            # per slot: {key: (pinned, dev)}
        self.ready = [torch.cuda.Event() for _ in range(depth)]     # H2D of the slot finished
        self.free = [None] * depth             # consumer finished with the slot (event on the consumer's stream)
        self.q = queue.Queue(maxsize=depth)
        self.free_q = queue.Queue()
        for i in range(depth):
            self.free_q.put(i)
        self._last = None
        self.h2d_bytes = 0
        self.thread = threading.Thread(target=self._produce, daemon=True)
        self.thread.start()

    def _buffers(self, slot, key, t):
        ent = self.slots[slot] or {}
        self.slots[slot] = ent
        dt = self.dtype if (self.dtype is not None and t.is_floating_point()) else t.dtype
        cur = ent.get(key)
        if cur is None or cur[1].shape != t.shape or cur[1].dtype != dt:
            cur = [None, torch.empty(t.shape, dtype=dt, device=self.device)]
            ent[key] = cur
        if cur[0] is None and not (t.is_pinned() and t.dtype == dt):
            # page-locked staging only for sources that are not page-locked themselves (cudaHostAlloc is slow and
            # synchronises the device: it happens once per slot, on the first batches)
            cur[0] = torch.empty(t.shape, dtype=dt).pin_memory()
        return cur

    def _produce(self):
        try:
            torch.cuda.set_device(self.device)
            for batch in self.src:
                slot = self.free_q.get()
                if slot is None:
                    return
                if self.free[slot] is not None:
                    self.ready[slot].synchronize()     # the slot's previous H2D has left its pinned buffers
                    self.copy_stream.wait_event(self.free[slot])
                parts = batch if isinstance(batch, (tuple, list)) else (batch,)
                outs = []
                with torch.cuda.stream(self.copy_stream):
                    for pi, part in enumerate(parts):
                        out = {}
                        for k, v in part.items():
                            if torch.is_tensor(v):
                                pinned, dev = self._buffers(slot, (pi, k), v)
                                if v.is_pinned() and v.dtype == dev.dtype:
                                    src = v                      # already page-locked: copy from where it lies
                                else:
                                    pinned.copy_(v)              # (dtype conversion, if any, happens here on the host)
                                    src = pinned
                                dev.copy_(src, non_blocking=True)
                                self.h2d_bytes += src.numel() * src.element_size()
                                out[k] = dev
                            else:
                                out[k] = v
                        outs.append(out)
                    self.ready[slot].record(self.copy_stream)
                self.q.put((slot, tuple(outs) if isinstance(batch, (tuple, list)) else outs[0]))
            self.q.put(None)
        except Exception as e:  # surface producer errors in the consumer
            self.q.put(e)

    def __iter__(self):
        return self

    def __next__(self):
        self._release()
        item = self.q.get()
        if item is None:
            raise StopIteration
        if isinstance(item, Exception):
            raise item
        slot, out = item
        torch.cuda.current_stream().wait_event(self.ready[slot])
        self._last = slot
        return out

    def _release(self):
        """The previous batch's device slot may be overwritten once the consumer's stream reaches this point."""
        if self._last is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.free[self._last] = ev
            self.free_q.put(self._last)
            self._last = None

    def close(self):
        self._release()
        self.free_q.put(None)
