"""Double-buffered host->device staging: while step i runs, the tensors of step i+1 are copied from pinned host memory
on a side stream (replaces the 8 blocking `.to(device)` calls of models/train_origin.py:95-104)."""
import torch


class DevicePrefetcher:
    def __init__(self, loader, device, host_indices=()):
        """host_indices: positions of the batch tuple that stay on the host (e.g. txt_labels: the MLM row selection is
        host-side integer work, so shipping them would only add a device->host sync)."""
        self.loader, self.device, self.host_indices = loader, torch.device(device), set(host_indices)
        self.stream = torch.cuda.Stream(device=self.device)
        self.bytes_last = 0

    def _stage(self, item):
        with torch.cuda.stream(self.stream):
            out, n = [], 0
            for i, t in enumerate(item):
                if torch.is_tensor(t) and i not in self.host_indices:
                    n += t.numel() * t.element_size()
                    out.append(t.to(self.device, non_blocking=True))
                else:
                    out.append(t)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev, n

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev, n = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in cur:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(torch.cuda.current_stream(self.device))
            self.bytes_last = n
            yield cur

    def __len__(self):
        return len(self.loader)
