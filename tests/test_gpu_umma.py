"""tcgen05 primitives (descriptor encodings, TMEM, commit/mbarrier) pinned on one 128 x 128 tile."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [32, 128, 192])
@pytest.mark.parametrize("ts_mode", [0, 1])
@pytest.mark.parametrize("passes", [1, 3])
def test_umma_tile_f16(K, ts_mode, passes):
    """kind::f16 with fp16 (hi, lo) pairs: 16-bit core-matrix layout, packed TMEM A operand, K = 16 per instruction."""
    from packppi_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(100 + K + 7 * ts_mode + passes)
    A = torch.randn(128, K, generator=g).to(dev)
    W = (torch.randn(128, K, generator=g) * 0.1).to(dev)
    D = torch.zeros(128, 128, device=dev)
    _lib.call("pp_selftest_umma_f16", A, W, D, K, passes, ts_mode)
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t()).float()
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    # plain fp16 keeps 11 bits of each input; the rounded (hi, lo) pairs keep 22
    assert err < (2e-6 if passes == 3 else 3e-3) * max(scale, 1.0) * (K / 32) ** 0.5, (err, scale)


def test_tma_gather4_semantics():
    """cp.async.bulk.tensor.2d ... tile::gather4 with a tensor-map box of {32 floats, 1 row} and 128-byte swizzle:
    four arbitrary rows land as four consecutive 128-byte rows, 16-byte unit u of shared-memory row r at
    r * 128 + ((u ^ (r & 7)) << 4) - the layout the staged tiles of csrc/mpnn_tc.cu use."""
    from packppi_b200 import _lib
    dev = torch.device("cuda:0")
    src = torch.arange(64 * 128, dtype=torch.float32, device=dev).reshape(64, 128)
    out = torch.zeros(256, device=dev)
    rows, col = (5, 17, 2, 9), 32
    _lib.call("pp_selftest_gather4", src, 64, 1, col, *rows, out)
    torch.cuda.synchronize()
    o = out.cpu().reshape(8, 8, 4)
    for r, srow in enumerate(rows):
        for u in range(8):
            want = src[srow, col + 4 * u:col + 4 * u + 4].cpu()
            assert torch.equal(o[r, u ^ (r & 7)], want), (r, u)
    assert bool((o[4:] == -1).all())


def test_tensor_core_accumulation_truncates_and_promotion_fixes_it():
    """The finding behind csrc/node_post_tc.cu: accumulating a long all-positive sum in TMEM is biased low (the fp32
    accumulator truncates), starting a fresh accumulator every K = 16 and summing in fp32 registers is not."""
    from packppi_b200 import _lib
    dev = torch.device("cuda:0")
    K = 512
    g = torch.Generator().manual_seed(K)
    A = (torch.rand(128, K, generator=g) + 0.5).to(dev)
    W = (torch.rand(128, K, generator=g) + 0.5).to(dev)
    ref = A.double() @ W.double().t()
    bias = {}
    for name, ts in (("tmem", 0), ("promoted", 2)):
        D = torch.zeros(128, 128, device=dev)
        _lib.call("pp_selftest_umma_f16", A, W, D, K, 3, ts)
        torch.cuda.synchronize()
        bias[name] = ((D.double() - ref) / ref).mean().item()
    assert bias["tmem"] < -2e-6, bias          # measured -3.4e-6 at K = 512 (-7e-7 at K = 128)
    assert abs(bias["promoted"]) < 3e-7, bias  # measured -1.0e-7 (the 22-bit operand split)
