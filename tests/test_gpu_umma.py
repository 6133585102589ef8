"""tcgen05 primitives (descriptor encodings, TMEM, commit/mbarrier) pinned on one 128 x 128 tile."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [32, 128, 192])
@pytest.mark.parametrize("ts_mode", [0, 1])
@pytest.mark.parametrize("passes", [1, 3])
def test_umma_tile(K, ts_mode, passes):
    from packppi_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(K + 7 * ts_mode + passes)
    A = torch.randn(128, K, generator=g).to(dev)
    W = (torch.randn(128, K, generator=g) * 0.1).to(dev)
    D = torch.zeros(128, 128, device=dev)
    _lib.call("pp_selftest_umma", A, W, D, K, passes, ts_mode)
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t()).float()
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    # plain TF32 keeps 11 bits of each input; the 3-pass split keeps 21+
    assert err < (3e-6 if passes == 3 else 3e-3) * max(scale, 1.0) * (K / 32) ** 0.5, (err, scale)


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
