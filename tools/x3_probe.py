"""Can three tf32 tensor-core products reach fp32-grade accuracy here?  Emulates the hi/lo operand split with three
calls of the library's tf32 GEMM (each accumulating over the full K inside the tensor core) and measures, against fp64,
the error of z = x W^T on the shared layer's and the TRN's shapes, and how many ReLU units would flip.
    python tools/x3_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ta3n_b200  # noqa: E402
from ta3n_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)


def gemm(A, B, engine):
    ta3n_b200.set_gemm_engine(engine)
    C = torch.empty(A.shape[0], B.shape[0], device=dev)
    _lib.check(lib.ta3n_gemm_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), A.shape[0], B.shape[0], A.shape[1],
                                torch.cuda.current_stream().cuda_stream))
    return C


def tf32_rn(x):      # round to nearest tf32 (10-bit mantissa), as the TFLOAT32 tensor map does
    i = x.view(torch.int32)
    r = ((i + 0x00000FFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)


for (M, N, K, wscale, relu_in) in [(2560, 512, 2048, 0.001, False), (512, 256, 2560, 0.02, True), (512, 256, 256, 0.06, True)]:
    A = torch.randn(M, K, generator=g)
    if relu_in:
        A = A.clamp_min(0) * 0.05
    B = torch.randn(N, K, generator=g) * wscale
    bias = torch.randn(N, generator=g) * wscale
    A, B, bias = A.to(dev), B.to(dev), bias.to(dev)
    ref = A.double() @ B.double().t() + bias.double()
    sig = ref.std().item()

    def report(tag, C):
        z = C.double() + bias.double()
        err = (z - ref)
        flips = ((z > 0) != (ref > 0)).sum().item()
        print(f"  {tag:14s} normwise {err.norm().item() / ref.norm().item():.2e}  rms err / sigma {err.std().item() / sig:.2e}  "
              f"mean err / sigma {err.mean().item() / sig:+.2e}  sign flips {flips} of {ref.numel()}")

    print(f"M={M} N={N} K={K}")
    report("fp32 simt", gemm(A, B, "fp32"))
    report("tf32", gemm(A, B, "tf32"))
    Ah, Bh = tf32_rn(A), tf32_rn(B)
    Al, Bl = A - Ah, B - Bh
    C3 = gemm(Ah, Bh, "tf32") + gemm(Al, Bh, "tf32") + gemm(Ah, Bl, "tf32")
    report("tf32 x3 (RN)", C3)
    C4 = C3 + gemm(Al, Bl, "tf32")
    report("tf32 x4 (RN)", C4)
    # K chunked: partial accumulation of 256 columns at a time, summed in fp32 outside the tensor core
    Cc = torch.zeros(M, N, device=dev)
    for k0 in range(0, K, 256):
        s = slice(k0, min(K, k0 + 256))
        Cc += gemm(Ah[:, s].contiguous(), Bh[:, s].contiguous(), "tf32") + gemm(Al[:, s].contiguous(), Bh[:, s].contiguous(), "tf32") + \
            gemm(Ah[:, s].contiguous(), Bl[:, s].contiguous(), "tf32")
    report("x3, K by 256", Cc)
    report("engine tf32x3", gemm(A, B, "tf32x3"))
    report("torch fp32", (A @ B.t()))
