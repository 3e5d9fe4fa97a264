"""Public host-facing API of the auto-labeling hot path: tracks in, refined 7-DoF boxes out.

``StaticAutoLabeler.label_device`` is the resident-tensor step (segmentation -> gather -> box head ->
decode); ``label_host`` takes pinned HOST buffers, streams them to the GPU in chunks on a copy stream
(double-buffered, overlapped with compute) and returns the boxes in pinned host memory -- this is what
the reference's eval loop does per DataLoader batch (tools/static_eval.py:256-290: H2D of pts /
init_box, model forward, argmax + class2angle / class2size decode, D2H).
"""
import torch

from . import ops


class StaticAutoLabeler:
    def __init__(self, model, chunk_tracks=2048, first_chunk_tracks=None):
        self.model = model
        self.chunk = int(chunk_tracks)
        # the first chunk's H2D copy is the only one not hidden behind compute: keep it small
        self.first = int(first_chunk_tracks) if first_chunk_tracks else max(1, self.chunk // 4)
        self._staging = None

    def _schedule(self, T):
        """Chunk boundaries: a short first chunk, then full chunks."""
        bounds, t0 = [], 0
        size = min(self.first, self.chunk)
        while t0 < T:
            t1 = min(T, t0 + size)
            bounds.append((t0, t1))
            t0, size = t1, self.chunk
        return bounds

    @torch.no_grad()
    def label_device(self, pts, init_box, bbox_gt=None):
        """pts (bs,3,n) CUDA (any strides), init_box (bs,7) -> boxes (bs,7) f32 [centre, l, w, h, heading]."""
        # the decode (tools/static_eval.py:270-288) runs in the epilogue of the fused head kernel
        _, boxes = self.model.forward_boxes(pts, init_box, bbox_gt if bbox_gt is not None else init_box)
        return boxes

    def _buffers(self, n, dev):
        key = (n, str(dev))
        if self._staging is None or self._staging[0] != key:
            bufs = [(torch.empty((self.chunk, n, 3), device=dev, dtype=torch.float32),
                     torch.empty((self.chunk, 7), device=dev, dtype=torch.float32)) for _ in range(2)]
            self._staging = (key, bufs, torch.cuda.Stream(device=dev))
        return self._staging[1], self._staging[2]

    @torch.no_grad()
    def label_host(self, pts_host, init_box_host, out_host=None):
        """pts_host (T,n,3) pinned f32 point-major, init_box_host (T,7) pinned f32 -> (T,7) pinned f32."""
        assert pts_host.is_pinned() and init_box_host.is_pinned(), "host buffers must be pinned"
        T, n, _ = pts_host.shape
        dev = next(self.model.parameters()).device
        if out_host is None:
            out_host = torch.empty((T, 7), dtype=torch.float32).pin_memory()
        bufs, copy_stream = self._buffers(n, dev)
        main = torch.cuda.current_stream(dev)
        free_ev = [None, None]
        for ci, (t0, t1) in enumerate(self._schedule(T)):
            k = t1 - t0
            dp, db = bufs[ci & 1]
            with torch.cuda.stream(copy_stream):
                if free_ev[ci & 1] is not None:
                    copy_stream.wait_event(free_ev[ci & 1])
                dp[:k].copy_(pts_host[t0:t1], non_blocking=True)
                db[:k].copy_(init_box_host[t0:t1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            main.wait_event(ready)
            boxes = self.label_device(dp[:k].transpose(2, 1), db[:k])
            out_host[t0:t1].copy_(boxes, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            free_ev[ci & 1] = ev
        main.synchronize()
        if getattr(self.model, "precision", "fp32") != "fp32":
            from . import engine_bf16
            engine_bf16.check_abort("StaticAutoLabeler.label_host", dev)
        return out_host
