"""Host -> device feeding of the hot path (reference: tools/utils.py:44-52 `to_cuda`, a blocking
`.cuda()` per tensor on the compute stream).

DevicePrefetcher copies the NEXT batch from pinned host memory on its own CUDA stream while the current batch is
being processed, and (optionally) expands 8-bit RGB + label maps into the fp32 `input` tensor on the device
(`pack_input`), so that only the bytes the dataset really holds cross PCIe."""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional

import torch


class DevicePrefetcher:
    """Iterate over device-resident batches, one batch ahead of the consumer.

    `batches` yields dicts of pinned host tensors.  If a batch holds "rgb" (uint8 or fp32) and "label" (uint8) and
    `num_lyt` is given, the yielded dict gets "input" = pack_input(rgb, label).  Tensors of a yielded batch stay valid
    until the consumer asks for the batch after the next one (two device buffers per key)."""

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device, num_lyt: Optional[int] = None, depth: int = 2,
                 input_dtype=torch.float32):
        self.it: Iterator = iter(batches)
        self.device = torch.device(device)
        self.num_lyt = num_lyt
        self.input_dtype = input_dtype                    # torch.bfloat16: input of the bf16-storage inference variant
        self.stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.slots = [dict() for _ in range(depth)]      # reusable device buffers
        self.ready = [None] * depth                       # event: copy (+ packing) of the slot finished
        self.freed = [None] * depth                       # event: consumer finished with the slot
        self.head = 0
        self.last = None                                  # slot of the batch the consumer currently holds
        self.queue = []
        self._fill()

    def _fill(self):
        while len(self.queue) < self.depth - 1:           # one slot belongs to the consumer, the others run ahead
            try:
                host = next(self.it)
            except StopIteration:
                return
            s = self.head
            self.head = (self.head + 1) % self.depth
            if self.freed[s] is not None:
                self.stream.wait_event(self.freed[s])
            from . import functional as Fn
            with torch.cuda.stream(self.stream):
                dev = self.slots[s]
                out = {}
                for k, t in host.items():
                    buf = dev.get(k)
                    if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                        buf = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                        dev[k] = buf
                    buf.copy_(t, non_blocking=True)
                    out[k] = buf
                if self.num_lyt is not None and "rgb" in out and "label" in out:
                    inp = dev.get("input")
                    B, T, _, Hd, Wd = out["rgb"].shape
                    if inp is None or inp.shape != (B, T, 3 + self.num_lyt, Hd, Wd):
                        inp = torch.empty(B, T, 3 + self.num_lyt, Hd, Wd, device=self.device, dtype=self.input_dtype)
                        dev["input"] = inp
                    out["input"] = Fn.pack_input(out["rgb"], out["label"], self.num_lyt, out=inp, dtype=self.input_dtype)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.ready[s] = ev
            self.queue.append((s, out))

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, torch.Tensor]:
        if not self.queue:
            raise StopIteration
        s, out = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        # asking for a new batch releases the previous one: its slot may be overwritten once everything enqueued on
        # the consumer's stream so far has run
        if self.last is not None:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.freed[self.last] = ev
        self.last = s
        self._fill()
        return out
