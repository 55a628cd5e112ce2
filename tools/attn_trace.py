"""clock64 trace of the attention kernel's max-free pass (build with -DATTN_TRACE: vittf_b200/csrc/build_variant.sh <name> -DATTN_TRACE,
run with VITTF_LIB=vittf_b200/libvittf_b200_<name>.so).  Per key block j of CTA 1: role 0 / 1 = softmax warps of query tile A / B
(loop top, scores arrived, exponential section begins, P stored), role 2 = tile A's MMA issuer."""
import os, sys, torch
sys.path.insert(0, '.')
from vittf_b200 import ops
B, tokens, heads = 32, 4097, int(os.environ.get("HEADS", 12))
D = heads * 64
qk = torch.randn(B * tokens, 2, D, device="cuda")
qk[:, 0] *= 0.125 * 1.4426950408889634
qk = qk.view(B * tokens, 2 * D).bfloat16()
vt = torch.randn(B * D, ops.tok_pad_of(tokens), device="cuda").bfloat16()
for _ in range(3): ops.attention_prescaled(qk, vt, B, tokens, heads, ops.tok_pad_of(tokens))
torch.cuda.synchronize()
os.environ["VITTF_ATTN_TRACE_DUMP"] = "1"
ops.attention_prescaled(qk, vt, B, tokens, heads, ops.tok_pad_of(tokens))
torch.cuda.synchronize()
