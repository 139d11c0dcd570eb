"""GPU: calls issued on different CUDA streams must not share scratch (one engine handle per stream)."""
import numpy as np
import pytest
import torch

import prysm_oracle as O
from conftest import rel_linf

pytestmark = pytest.mark.gpu


def test_two_streams_do_not_share_scratch():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import prysm_b200 as pb
    from prysm_b200 import propagation as P, _capi
    rng = np.random.default_rng(0)
    n = 512                                     # tuned focus path: uses the per-handle intermediate
    a = [(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64) for _ in range(2)]
    d = [pb.asdevice(x) for x in a]
    streams = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    outs = [None, None]
    for rep in range(8):                        # interleave launches so the two streams overlap on the device
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                outs[i] = P.focus(d[i], 2)
    torch.cuda.synchronize()
    for i in range(2):
        assert rel_linf(outs[i].cpu().numpy(), O.focus(a[i].astype(np.complex128), 2)) < 1e-6
    dev = torch.cuda.current_device()
    keys = [k for k in _capi._handles if k[0] == dev]
    assert len({k[1] for k in keys}) >= 2       # distinct handles for distinct streams
