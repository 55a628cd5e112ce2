"""Kernel micro-benchmarks (CUDA events, L2 flushed between iterations).  Diagnostic tool, not bench.py."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, tokens = 8, 4097
    for D, heads in ((384, 6), (768, 12)):
        M = B * tokens
        for name, N, K, epi in (("qkv", 3 * D, D, _lib.EPI_QKV_SPLIT), ("proj", D, D, _lib.EPI_BIAS_RESID_F32),
                                ("fc1", 4 * D, D, _lib.EPI_BIAS_GELU_BF16), ("fc2", D, 4 * D, _lib.EPI_BIAS_RESID_F32)):
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
            tf = 2.0 * M * N * K / (med * 1e-3) / 1e12
            ref_med, _ = timeit(lambda: torch.matmul(a, w.t()), flush=flush)
            res[f"gemm_{name}_D{D}"] = {"ms": med, "best_ms": best, "tflops": tf, "cublas_ms": ref_med,
                                       "cublas_tflops": 2.0 * M * N * K / (ref_med * 1e-3) / 1e12}
            print(f"gemm {name} D={D}: {med:.3f} ms {tf:.0f} TF/s (cuBLAS {ref_med:.3f} ms)", flush=True)
        qk = torch.randn(M, 2 * D, device="cuda").bfloat16()
        vt = torch.randn(B * D, ops.tok_pad_of(tokens), device="cuda").bfloat16()
        med, best = timeit(lambda: ops.attention(qk, vt, B, tokens, heads, ops.tok_pad_of(tokens)), flush=flush)
        fl = 4.0 * B * heads * tokens * tokens * 64
        res[f"attention_D{D}"] = {"ms": med, "best_ms": best, "tflops": fl / (med * 1e-3) / 1e12}
        print(f"attention D={D}: {med:.3f} ms {fl / (med * 1e-3) / 1e12:.0f} TF/s", flush=True)
        q = qk.view(B, tokens, 2, heads, 64)
        qq, kk = q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2)
        vv = vt.view(B, heads, 64, -1)[..., :tokens].transpose(2, 3).contiguous()
        med, _ = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv), flush=flush)
        res[f"sdpa_D{D}"] = {"ms": med, "tflops": fl / (med * 1e-3) / 1e12}
        print(f"torch sdpa D={D}: {med:.3f} ms {fl / (med * 1e-3) / 1e12:.0f} TF/s", flush=True)
    # similarity passes (cfg4-like, reduced): 384 x 64^3 fp16 -> 256^3, A = 8 / 32
    from vittf_b200.similarity import similarity_maps
    feats = torch.randn(384, 64, 64, 64, device="cuda").half()
    for A in (8, 32):
        protos = torch.nn.functional.normalize(torch.randn(A, 384, device="cuda"), dim=-1)
        offs = torch.arange(0, A + 1, A // 8, dtype=torch.int32, device="cuda")
        med1, _ = timeit(lambda: ops.sim_lowres(feats, protos), flush=flush)
        low = ops.sim_lowres(feats, protos)
        out = torch.empty(8, 256, 256, 256, device="cuda")
        med2, _ = timeit(lambda: ops.sim_upsample(low[0], low[1], (64, 64, 64), offs, (256, 256, 256), 0, out=out, layout=low[2]), flush=flush)
        bytes1 = feats.numel() * 2 + (A + 14) * 64 ** 3 * 4
        bytes2 = (A + 14) * 64 ** 3 * 4 + out.numel() * 4
        res[f"sim_A{A}"] = {"lowres_ms": med1, "lowres_gbs": bytes1 / med1 / 1e6, "upsample_ms": med2,
                           "upsample_gbs": bytes2 / med2 / 1e6, "gvox_s": 256 ** 3 / ((med1 + med2) * 1e-3) / 1e9}
        print(f"sim A={A}: lowres {med1:.3f} ms ({bytes1 / med1 / 1e6:.0f} GB/s), upsample {med2:.3f} ms ({bytes2 / med2 / 1e6:.0f} GB/s)", flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/microbench.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
