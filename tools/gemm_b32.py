"""GEMM timings at the bench's slice batch (B=32 images of 4097 tokens), CUDA events, L2 flushed."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import _lib, ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

B, tokens = int(os.environ.get("B", 32)), 4097
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for D in (384, 768):
    M = B * tokens
    for name, N, K, epi in (("qkv", 3 * D, D, _lib.EPI_QKV_SPLIT), ("proj", D, D, _lib.EPI_BIAS_RESID_F32),
                            ("fc1", 4 * D, D, _lib.EPI_BIAS_GELU_BF16), ("fc1 without GELU", 4 * D, D, _lib.EPI_BIAS_BF16),
                            ("fc2", D, 4 * D, _lib.EPI_BIAS_RESID_F32)):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bias = torch.zeros(N, device="cuda")
        tok_pad = ops.tok_pad_of(tokens)
        if epi == _lib.EPI_QKV_SPLIT:
            o1 = torch.empty(M, 2 * D, dtype=torch.bfloat16, device="cuda")
            o2 = torch.zeros(B * D, tok_pad, dtype=torch.bfloat16, device="cuda")
        elif epi == _lib.EPI_BIAS_RESID_F32:
            o1, o2 = torch.zeros(M, N, device="cuda"), None
        else:
            o1, o2 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), None
        med, best = timeit(lambda: ops.gemm_bf16(a, w, bias, epi, out=o1, out2=o2, tokens=tokens, tok_pad=tok_pad), flush=flush)
        ref, _ = timeit(lambda: torch.matmul(a, w.t()), flush=flush)
        print(f"B={B} D={D} gemm {name}: {med * 1e3:.1f} us {2.0 * M * N * K / (med * 1e-3) / 1e12:.0f} TF/s "
              f"(cuBLAS plain {2.0 * M * N * K / (ref * 1e-3) / 1e12:.0f} TF/s) ares_off={'VITTF_GEMM_NO_ARES' in os.environ}", flush=True)
