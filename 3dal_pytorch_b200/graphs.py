"""CUDA-graph replay of the eval forward for fixed input shapes.

The reference's eval loops call ``model(pts, box, gt)`` once per batch of 32 / 64 tracks (tools/static_eval.py:50-70,
tools/dynamic_eval.py:40-60).  At those sizes the forward is a sequence of ~10-25 short kernels and the time is launch
latency, not arithmetic (static one-box, 32 tracks x 4096 points: 0.63 ms eager).  ``GraphedForward`` captures that sequence
once -- the models' eval forward has no host synchronisation with the default gather policy -- and replays it as ONE
graph launch.  Same kernels, same order, same memory: the outputs are bit-identical to the eager call
(tests/test_gpu_graphs.py)."""
import torch

from . import engine_bf16


class GraphedForward:
    def __init__(self, model, *example_inputs, warmup=2):
        if model.training:
            raise ValueError("GraphedForward captures the eval forward; call model.eval() first")
        if getattr(model, "gather_policy", "strided") != "strided":
            raise ValueError("gather_policy %r needs a host round trip and cannot be captured" % model.gather_policy)
        self.model = model
        self.static_in = [t.clone() if torch.is_tensor(t) else t for t in example_inputs]
        self.device = next(t for t in self.static_in if torch.is_tensor(t)).device
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                     # weight packing, lazy allocations, the watchdog block: before the capture
                model(*self.static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        engine_bf16.check_abort("graph warm-up", self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = model(*self.static_in)

    def __call__(self, *inputs):
        """Copies the inputs into the captured buffers, replays, and returns the captured output dict -- its tensors are
        overwritten by the next call (clone what must survive)."""
        if len(inputs) != len(self.static_in):
            raise ValueError("expected %d inputs" % len(self.static_in))
        for dst, src in zip(self.static_in, inputs):
            if torch.is_tensor(dst):
                if src.shape != dst.shape:
                    raise ValueError("GraphedForward was captured for shape %s, got %s" % (tuple(dst.shape), tuple(src.shape)))
                dst.copy_(src, non_blocking=True)
        engine_bf16.check_abort("graph replay", self.device)
        self.graph.replay()
        return self.static_out
