"""CUDA-graph capture of a launch-bound composition.

Small problems on this path (a 6-level vortex stack on a 384^2 pupil is ~65 kernels of a few microseconds each)
are bound by launch latency, not by HBM or the tensor cores.  The engine's answer is a CUDA graph, not a tracing
compiler: run the composition once on a side stream so that every handle-owned resource exists (twiddle / chirp
tables, scratch arenas, kernel attributes), capture the same calls into a graph, and replay the graph per call.

    step = capture(lambda w: P.to_fpm_and_back_multiresolution(w, fpm, mex), pupil)
    out = step(new_pupil)          # copies new_pupil into the captured input buffer, one graph launch

Capturable: every asynchronous entry point of the C ABI.  NOT capturable: the calls that return host scalars
(`pb_dot`, `pb_mode_projection`, `pb_encircled_energy*`, `pb_moments` -- they synchronise the stream) and plan
construction from host coordinates (`prepare_executor`, `prepare_multiresolution` -- build those outside).
"""
import torch

from . import _capi


class CapturedGraph:
    """A replayable capture of fn(*inputs).  Tensor inputs are copied into static buffers owned by the capture;
    the returned tensors are static too (valid until the next call)."""

    def __init__(self, fn, *example_inputs, warmup=2):
        self._static_in = [x.detach().clone() if isinstance(x, torch.Tensor) else x for x in example_inputs]
        dev = next((x.device for x in self._static_in if isinstance(x, torch.Tensor)), None)
        if dev is None or dev.type != 'cuda':
            raise ValueError('capture() needs at least one CUDA tensor input')
        self._stream = torch.cuda.Stream(device=dev)
        self._key = (dev.index if dev.index is not None else torch.cuda.current_device(), self._stream.cuda_stream)
        _capi.pin_handle(*self._key)     # the graph replays into this stream's handle (tables + scratch): keep it alive
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):          # warm-up on the capture stream: its handle builds its tables here
            for _ in range(max(1, warmup)):
                fn(*self._static_in)
        torch.cuda.current_stream(dev).wait_stream(self._stream)
        self._stream.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=self._stream):
            self._static_out = fn(*self._static_in)

    def __call__(self, *inputs):
        if len(inputs) != len(self._static_in):
            raise TypeError(f'captured with {len(self._static_in)} inputs, called with {len(inputs)}')
        for buf, x in zip(self._static_in, inputs):
            if isinstance(buf, torch.Tensor):
                if x is not buf:
                    buf.copy_(x, non_blocking=True)
            elif x != buf:
                raise ValueError('non-tensor arguments are baked into the capture and cannot change')
        self._graph.replay()
        return self._static_out

    def close(self):
        """Release the graph and the engine state of its capture stream."""
        self._graph = None
        self._static_out = None
        _capi.release_handle(*self._key)

    @property
    def inputs(self):
        """The static input buffers (write into them directly to skip the per-call copy)."""
        return self._static_in


def capture(fn, *example_inputs, warmup=2):
    """Capture fn(*example_inputs) into a CUDA graph; returns a callable with the same signature."""
    return CapturedGraph(fn, *example_inputs, warmup=warmup)
