"""The gradient collective's kernel (csrc/peer_reduce.cu) on ONE GPU: a world of one rank runs the same code - epoch
flags, last-CTA detection, slice loop, fused scale - against its own buffer.  The multi-rank behaviour (unicast and
in-switch reduction, agreement with NCCL to the last bit) is checked by `bench.py --gpus N` in every run (`collective` key)
and recorded in profiles/r2_peer_reduce_probe.txt; the host-side ordering logic by tests/test_dist_gloo.py."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _call(buf, flags, slot, offset, n, scale, ctas, world=1, rank=0):
    from neuradar_b200 import _lib

    bufs = (C.c_uint64 * 1)(buf.data_ptr())
    sigs = (C.c_uint64 * 1)(flags.data_ptr())
    _lib.call("nrb_peer_all_reduce", bufs, sigs, 0, rank, world, slot, offset, n, float(scale), ctas, _lib.stream_ptr())


@pytest.mark.parametrize("ctas", [0, 1, 8])
def test_single_rank_world_scales_in_place_and_can_be_repeated(ctas):
    n = (1 << 20) + 4 * 37  # not a multiple of the block size
    g = torch.Generator().manual_seed(3)
    x = torch.randn((n + 8,), generator=g).to(DEV)
    want = x.clone()
    flags = torch.zeros((256,), dtype=torch.int32, device=DEV)
    for call in range(4):  # the epoch advances, the CTA counter returns to zero: the call is repeatable
        _call(x, flags, slot=1, offset=4, n=n, scale=0.5, ctas=ctas)
        want[4: 4 + n] *= 0.5
        torch.cuda.synchronize()
        assert torch.equal(x, want), call
    f = flags.cpu()
    assert int(f[64 + 32]) == 4 and int(f[64 + 33]) == 0          # slot 1: epoch word, CTA counter
    assert int(f[64 + 0]) == 4 and int(f[64 + 16]) == 4            # this rank's barrier flags of both phases
    assert int(f[:64].abs().sum()) == 0 and int(f[128:].abs().sum()) == 0  # other slots untouched


def test_graph_replay():
    n = 1 << 18
    x = torch.ones((n,), device=DEV)
    flags = torch.zeros((256,), dtype=torch.int32, device=DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        _call(x, flags, 0, 0, n, 2.0, 8)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        _call(x, flags, 0, 0, n, 2.0, 8)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(x, torch.full((n,), 2.0 ** 4, device=DEV))  # one eager call + three replays (capturing runs nothing)


def test_argument_checks():
    from neuradar_b200 import _lib

    x = torch.ones((64,), device=DEV)
    flags = torch.zeros((256,), dtype=torch.int32, device=DEV)
    with pytest.raises(_lib.NeuradarB200Error):
        _call(x, flags, 0, 2, 32, 1.0, 0)        # offset not a multiple of 4 floats
    with pytest.raises(_lib.NeuradarB200Error):
        _call(x, flags, 7, 0, 32, 1.0, 0)        # no such flag slot
    with pytest.raises(_lib.NeuradarB200Error):
        _call(x, flags, 0, 0, 32, 1.0, 0, world=9)  # more ranks than one NVSwitch domain
    _call(x, flags, 0, 0, 0, 1.0, 0)             # nothing to do
    torch.cuda.synchronize()
    assert torch.equal(x, torch.ones((64,), device=DEV))
