"""Hardware-semantics probes the window-convolution kernels rely on (csrc/winconv.cuh):
row-shifted UMMA descriptors over a 128-byte-swizzled buffer, K-major and MN-major, and two overlapping MN-major atoms
(paired kernel taps) in one M = 128 instruction.  If a future driver/GPU changes these, the conv layers are wrong."""
import pytest
import torch

from r2d2_b200 import _lib

pytestmark = pytest.mark.gpu


def _probe(A, B, shift, mode):
    D = torch.zeros(128, 32, device="cuda")
    _lib.check(_lib.lib().r2d2_debug_shift_probe(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), shift, mode, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return D


@pytest.fixture(scope="module")
def operands():
    _lib.require_device()
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(144, 64, device="cuda", generator=g).bfloat16().contiguous()
    B = torch.randn(32, 64, device="cuda", generator=g).bfloat16().contiguous()
    return A, B


@pytest.mark.parametrize("shift", [0, 1, 3, 7, 8, 11])
def test_row_shifted_k_major_descriptor(operands, shift):
    A, B = operands
    ref = A[shift:shift + 128].float() @ B.float().t()
    assert torch.allclose(_probe(A, B, shift, 0), ref, atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("shift", [0, 1, 5, 9])
def test_row_shifted_mn_major_descriptor(operands, shift):
    A, B = operands
    ref = A[shift:shift + 64].float().t() @ B.float().t()                    # M = 64: row m in TMEM lane (m & 15) + 32 (m >> 4)
    lanes = torch.tensor([(m & 15) + 32 * (m >> 4) for m in range(64)], device="cuda")
    assert torch.allclose(_probe(A, B, shift, 2)[lanes], ref, atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("dist", [1, 7, 9, 21])
def test_paired_taps_share_one_mma(operands, dist):
    A, B = operands
    s = 3
    ref = torch.cat([A[s:s + 64].float().t() @ B.float().t(), A[s + dist:s + dist + 64].float().t() @ B.float().t()])
    assert torch.allclose(_probe(A, B, s, 2 + dist), ref, atol=1e-3, rtol=1e-3)


def test_ss_mode_issue_rate_matches_operand_read_model():
    """cycles per M x N x 16 instruction = max(128 N / 256, (M + N) * 32 B / 128 B/clk) within 10 %."""
    _lib.require_device()
    out = torch.zeros(1, dtype=torch.int64, device="cuda")
    for M, N in [(128, 64), (128, 128), (128, 256), (64, 64)]:
        _lib.check(_lib.lib().r2d2_debug_mma_rate(M, N, 2000, 0, 1, _lib.ptr(out), _lib.stream_ptr()))
        torch.cuda.synchronize()
        clk = out.item() / 8000
        model = max(128 * N / 256, (M + N) / 4)
        assert abs(clk - model) <= 0.1 * model, (M, N, clk, model)


def test_tmem_a_operand_and_sw64_b_tiles():
    """tcgen05.mma with A in tensor memory (written by tcgen05.st) and B as K-major SWIZZLE_64B tiles (recurrence2.cuh)."""
    from r2d2_b200 import _lib
    _lib.require_device()
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(128, 64, device="cuda", generator=g).bfloat16().contiguous()
    B = torch.randn(16, 64, device="cuda", generator=g).bfloat16().contiguous()
    D = torch.full((128, 16), float("nan"), device="cuda")
    _lib.check(_lib.lib().r2d2_debug_ts_probe(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < 1e-4, err
