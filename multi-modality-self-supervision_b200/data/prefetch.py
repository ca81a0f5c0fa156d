"""Double-buffered host->device staging: while step i runs, the tensors of step i+1 are copied from pinned host memory
on a side stream (replaces the 8 blocking `.to(device)` calls of models/train_origin.py:95-104).

The device side is two PERSISTENT buffer sets that the copies land in (`copy_`, no allocation per step): allocating fresh
tensors on the side stream every step made the caching allocator grow its per-stream pool for many steps — each growth a
synchronous cudaMalloc that drains the GPU queue (seen as an end-to-end rate that sporadically halved on a cold start).
"""
import torch


class DevicePrefetcher:
    def __init__(self, loader, device, host_indices=()):
        """host_indices: positions of the batch tuple that stay on the host (e.g. txt_labels: the MLM row selection is
        host-side integer work, so shipping them would only add a device->host sync)."""
        self.loader, self.device, self.host_indices = loader, torch.device(device), set(host_indices)
        self.stream = torch.cuda.Stream(device=self.device)
        self.bytes_last = 0
        self._bufs = [None, None]            # two sets of device tensors, re-used every other step
        self._free = [None, None]            # event: the step that consumed set k has been enqueued (and will finish) before this

    def _stage(self, item, k):
        slot = k & 1
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])       # the consumer of this buffer set is done with it
            bufs = self._bufs[slot]
            if bufs is None or len(bufs) != len(item):
                bufs = [None] * len(item)
            out, n = [], 0
            for i, t in enumerate(item):
                if torch.is_tensor(t) and i not in self.host_indices:
                    n += t.numel() * t.element_size()
                    b = bufs[i]
                    if b is None or b.shape != t.shape or b.dtype != t.dtype:
                        b = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                        bufs[i] = b
                    b.copy_(t, non_blocking=True)
                    out.append(b)
                else:
                    out.append(t)
            self._bufs[slot] = bufs
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev, n

    def __iter__(self):
        it = iter(self.loader)
        k = 0
        try:
            nxt = self._stage(next(it), k)
        except StopIteration:
            return
        while nxt is not None:
            cur, ev, n = nxt
            try:
                nxt = self._stage(next(it), k + 1)
            except StopIteration:
                nxt = None
            cs = torch.cuda.current_stream(self.device)
            cs.wait_event(ev)
            self.bytes_last = n
            yield cur
            # the consumer has enqueued its work on the compute stream: buffer set k may be overwritten after this point
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self._free[k & 1] = done
            k += 1

    def __len__(self):
        return len(self.loader)
