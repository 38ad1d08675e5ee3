"""tcgen05 GEMM kernels (single-CTA cp.async path, CTA-pair TMA path) vs float64 matmul on plain matrices."""
import itertools
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

def _split(x):
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi.contiguous(), lo.contiguous()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 512), (304, 64, 576), (64, 2048, 512), (1000, 32, 256),
                                   (136, 528, 528), (128, 16, 2560), (2048, 512, 5440), (64, 576, 20000)])
def test_umma2_split_operands_matches_fp64(M, N, K):
    """v2 kernel: cp.async chunk producers, K-major and MN-major descriptors, bf16x3 and fast modes."""
    from r2d2_b200 import _lib
    _lib.require_device()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    lines = []
    for am, bm in itertools.product((0, 1), (0, 1)):
        A = torch.randn((M, K) if am == 0 else (K, M), device="cuda", generator=g)
        B = torch.randn((N, K) if bm == 0 else (K, N), device="cuda", generator=g)
        ah, al = _split(A)
        bh, bl = _split(B)
        ref = (A if am == 0 else A.t()).double() @ (B if bm == 0 else B.t()).double().t()
        scale = ref.abs().max().item()
        for ubn in (16, 32, 64, 128, 256):
            if bm == 1 and (ubn < 64):
                continue
            for fast, tol in ((0, 3e-5 if K < 8192 else 1e-4), (1, 2e-2)):
                for splits in ((1, 3) if K >= 512 else (1,)):
                    prev = _lib.lib().r2d2_set_fast_math(fast)
                    C = torch.full((splits, M, N), float("nan"), device="cuda")
                    _lib.check(_lib.lib().r2d2_debug_gemm2(ubn, am, bm, M, N, K, _lib.ptr(ah), _lib.ptr(al), _lib.ptr(bh),
                                                           _lib.ptr(bl), _lib.ptr(C), splits, _lib.stream_ptr()))
                    torch.cuda.synchronize()
                    _lib.lib().r2d2_set_fast_math(prev)
                    C = C.sum(0)
                    assert torch.isfinite(C).all(), (ubn, am, bm, fast, splits)
                    err = (C.double() - ref).abs().max().item() / scale
                    lines.append(f"v2 M{M} N{N} K{K} am{am} bm{bm} ubn{ubn} fast{fast} sp{splits}: rel err {err:.2e}")
                    assert err < tol, lines[-1]
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "gemm_diag.txt"), "a") as f:
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 512), (304, 264, 576), (136, 528, 528), (64, 2048, 512),
                                   (5440, 512, 3136), (2048, 528, 5440), (512, 3136, 2720), (1000, 3136, 512)])
def test_umma3_pair_gemm_matches_fp64(M, N, K):
    """v3 kernel: CTA pairs (cta_group::2), TMA producers, K-major and MN-major operands, split-K, ragged edges."""
    from r2d2_b200 import _lib
    _lib.require_device()
    g = torch.Generator(device="cuda").manual_seed(3 * M + N + K)
    lines = []
    for am, bm in itertools.product((0, 1), (0, 1)):
        A = torch.randn((M, K) if am == 0 else (K, M), device="cuda", generator=g)
        B = torch.randn((N, K) if bm == 0 else (K, N), device="cuda", generator=g)
        ah, al = _split(A)
        bh, bl = _split(B)
        ref = (A if am == 0 else A.t()).double() @ (B if bm == 0 else B.t()).double().t()
        scale = ref.abs().max().item()
        for fast, tol in ((0, 3e-5), (1, 2e-2)):
            for splits in ((1, 3) if K >= 512 else (1,)):
                prev = _lib.lib().r2d2_set_fast_math(fast)
                C = torch.zeros((splits, M, N), device="cuda")
                _lib.check(_lib.lib().r2d2_debug_gemm3(am, bm, M, N, K, _lib.ptr(ah), _lib.ptr(al), _lib.ptr(bh), _lib.ptr(bl),
                                                       _lib.ptr(C), splits, _lib.stream_ptr()))
                torch.cuda.synchronize()
                _lib.lib().r2d2_set_fast_math(prev)
                C = C.sum(0)
                assert torch.isfinite(C).all(), (am, bm, fast, splits)
                err = (C.double() - ref).abs().max().item() / scale
                lines.append(f"v3 M{M} N{N} K{K} am{am} bm{bm} fast{fast} sp{splits}: rel err {err:.2e}")
                assert err < tol, lines[-1]
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "gemm_diag.txt"), "a") as f:
        f.write("\n".join(lines) + "\n")
