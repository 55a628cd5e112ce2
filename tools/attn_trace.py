import os, sys, torch
sys.path.insert(0, '.')
from vittf_b200 import ops
B, tokens, heads = 8, 4097, 6
D = heads * 64
qk = torch.randn(B * tokens, 2 * D, device="cuda").bfloat16()
vt = torch.randn(B * D, ops.tok_pad_of(tokens), device="cuda").bfloat16()
for _ in range(3): ops.attention(qk, vt, B, tokens, heads, ops.tok_pad_of(tokens))
torch.cuda.synchronize()
os.environ["VITTF_ATTN_TRACE_DUMP"] = "1"
ops.attention(qk, vt, B, tokens, heads, ops.tok_pad_of(tokens))
