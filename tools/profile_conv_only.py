"""A few isolated conv launches (K25, 768->768, B=32, T=752) for `ncu --set full -k regex:tapgemm`."""
import sys
import torch
sys.path.insert(0, ".")
from openseq2seq_b200 import _lib as L
lib = L.load()
st = L.stream_ptr()
B, T, K, dil, Cin, Cout = 32, 752, 25, 1, 768, 768
padl = 12
x = torch.randn(B, T, Cin, device="cuda").bfloat16()
w = (torch.randn(K, Cin, Cout, device="cuda") / (K * Cin) ** 0.5).bfloat16()
y = torch.empty(B, T, Cout, device="cuda").half()
dy = torch.randn(B, T, Cout, device="cuda").bfloat16()
dx = torch.empty(B, T, Cin, device="cuda").bfloat16()
dw = torch.empty(K, Cin, Cout, device="cuda")
for _ in range(2):
    lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 3, None, st)
    lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, T, Cin, Cout, K, dil, padl, 0, st)
    lib.os2s_conv1d_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), B, T, Cin, Cout, K, dil, padl, st)
torch.cuda.synchronize()
print("done")
