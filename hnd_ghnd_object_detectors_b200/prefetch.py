"""Host <-> device plumbing of the training loop (torch streams only; no kernels of ours).

The reference copies every batch synchronously on the compute stream and reads `loss.item()` every
step (src/mimic_runner.py:44-45,56-58), which leaves the GPU idle during the PCIe copy and the host
idle during the step.  DevicePrefetcher issues the host->device copy of batch i+1 on a side stream
while step i runs; AsyncScalarReader moves the per-step loss to pinned host memory without blocking
and hands back the value of the PREVIOUS step, so one device->host read still happens per step.
"""
import torch


class DevicePrefetcher(object):
    """Wraps an iterable of (images, targets) with CPU tensors; yields them as CUDA tensors whose
    copies were enqueued one batch ahead on a dedicated copy stream."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _upload(self, batch):
        images, targets = batch
        with torch.cuda.stream(self.stream):
            imgs = [im if im.is_cuda else (im if im.is_pinned() else im.pin_memory()).to(
                self.device, non_blocking=True) for im in images]
            tgts = None if targets is None else [
                {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in t.items()}
                for t in targets]
        return imgs, tgts

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur_stream = torch.cuda.current_stream(self.device)
            cur_stream.wait_stream(self.stream)  # batch i is complete before the compute stream uses it
            imgs, tgts = nxt
            for t in imgs:
                t.record_stream(cur_stream)
            try:
                nxt = self._upload(next(it))  # batch i+1 overlaps step i
            except StopIteration:
                nxt = None
            yield imgs, tgts


class AsyncScalarReader(object):
    """push(t) enqueues a non-blocking copy of a 0-dim / 1-element CUDA tensor into pinned memory and
    returns the value pushed one call earlier (None the first time); flush() returns the last one."""

    def __init__(self, depth=2):
        self.bufs = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.n = 0

    def push(self, t):
        prev = self._read(self.n - 1) if self.n > 0 else None
        i = self.n % len(self.bufs)
        self.bufs[i].copy_(t.detach().reshape(1).float(), non_blocking=True)
        self.events[i].record()
        self.n += 1
        return prev

    def _read(self, k):
        i = k % len(self.bufs)
        self.events[i].synchronize()
        return float(self.bufs[i][0])

    def flush(self):
        return self._read(self.n - 1) if self.n > 0 else None
