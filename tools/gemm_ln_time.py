"""LayerNorm-folded GEMM chain against the classic chain (layernorm_kernel + reduce-add epilogue) at the bench's slice batch
(B images of 4097 tokens): CUDA events, L2 flushed.  Prints per-launch times and the per-block totals of both chains."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import _lib, ops  # noqa: E402


def timeit(fn, flush=None, reps=6, groups=3):
    """Per-call time of `reps` back-to-back launches between one event pair (median of `groups`): the Python wrapper's CPU
    time hides behind the previous launch instead of showing up as GPU idle time inside the measurement -- the operands
    of every call here (>= 0.8 GB) are far larger than the L2, so no flush is needed between calls."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(groups):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()                                   # un-timed: keeps the GPU busy while the timed launches are enqueued
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


B, tokens = int(os.environ.get("B", 64)), 4097
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for D in [int(d) for d in os.environ.get("DIMS", "768,384").split(",")]:
    M = B * tokens
    mp = ops.m_pad_of(M)
    tok_pad = ops.tok_pad_of(tokens)
    x = torch.randn(M, D, device="cuda")
    xt, xb, stats = ops.ln_prepare(x)
    ones, zeros = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    att = torch.randn(M, D, device="cuda").bfloat16()
    hid = torch.randn(M, 4 * D, device="cuda").bfloat16()
    w_qkv, w_proj = (torch.randn(3 * D, D, device="cuda") * 0.05).bfloat16(), (torch.randn(D, D, device="cuda") * 0.05).bfloat16()
    w_fc1, w_fc2 = (torch.randn(4 * D, D, device="cuda") * 0.05).bfloat16(), (torch.randn(D, 4 * D, device="cuda") * 0.02).bfloat16()
    b3, b1, b4 = torch.zeros(3 * D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(4 * D, device="cuda")
    cs3, cs4 = w_qkv.float().sum(1), w_fc1.float().sum(1)
    qk = torch.empty(M, 2 * D, dtype=torch.bfloat16, device="cuda")
    vt = torch.zeros(B * D, tok_pad, dtype=torch.bfloat16, device="cuda")
    hid_o = torch.empty(M, 4 * D, dtype=torch.bfloat16, device="cuda")
    xn = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    total = {"classic": 0.0, "folded": 0.0}

    def t(label, chain, fn, flops=0.0):
        med, _ = timeit(fn, flush=flush)
        total[chain] += med
        tf = f" {flops / (med * 1e-3) / 1e12:.0f} TF/s" if flops else ""
        print(f"B={B} D={D} {chain:8s} {label:28s} {med * 1e3:8.1f} us{tf}", flush=True)

    f_qkv, f_proj, f_fc = 2.0 * M * 3 * D * D, 2.0 * M * D * D, 2.0 * M * 4 * D * D
    t("layernorm x2", "classic", lambda: (ops.layernorm(x, ones, zeros), ops.layernorm(x, ones, zeros)))
    t("qkv", "classic", lambda: ops.gemm_bf16(xb, w_qkv, b3, _lib.EPI_QKV_SPLIT, out=qk, out2=vt, tokens=tokens, tok_pad=tok_pad), f_qkv)
    t("proj (reduce-add)", "classic", lambda: ops.gemm_bf16(att, w_proj, b1, _lib.EPI_BIAS_RESID_F32, out=x), f_proj)
    t("fc1 + GELU", "classic", lambda: ops.gemm_bf16(xb, w_fc1, b4, _lib.EPI_BIAS_GELU_BF16, out=hid_o), f_fc)
    t("fc2 (reduce-add)", "classic", lambda: ops.gemm_bf16(hid, w_fc2, b1, _lib.EPI_BIAS_RESID_F32, out=x), f_fc)
    t("qkv (LN in epilogue)", "folded", lambda: ops.gemm_bf16_ln(xb, w_qkv, b3, _lib.EPI_QKV_SPLIT, colsum=cs3, stats=stats, out=qk, out2=vt,
                                                               tokens=tokens, tok_pad=tok_pad), f_qkv)
    t("proj (stream + copy + sums)", "folded", lambda: ops.gemm_bf16_ln(att, w_proj, b1, _lib.EPI_BIAS_RESID_LN, xt=xt, out=xn), f_proj)
    t("fc1 + GELU (LN in epilogue)", "folded", lambda: ops.gemm_bf16_ln(xb, w_fc1, b4, _lib.EPI_BIAS_GELU_BF16, colsum=cs4, stats=stats, out=hid_o), f_fc)
    t("fc2 (stream + copy + sums)", "folded", lambda: ops.gemm_bf16_ln(hid, w_fc2, b1, _lib.EPI_BIAS_RESID_LN, xt=xt, out=xn), f_fc)
    print(f"B={B} D={D} per block: classic {total['classic'] * 1e3:.0f} us, folded {total['folded'] * 1e3:.0f} us "
          f"({100.0 * (1.0 - total['folded'] / total['classic']):.1f} % less)", flush=True)
    del x, xt, xb, att, hid, hid_o, qk, vt, xn
    torch.cuda.empty_cache()
