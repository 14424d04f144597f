"""Per-layer forward/backward deviation of JasperEngine against the storage-emulating oracle twin."""
import sys
import torch
sys.path.insert(0, ".")
from tests.common_cfg import MINI_JASPER
from tests.test_engine_gpu import _setup, _rel, _rel_l2
from oracle import torch_twin as TT

eng, params, feats, lens, labels, label_lens = _setup()
logits, out_lens = eng.forward(feats.cuda().bfloat16().contiguous(), lens.cuda())
torch.cuda.synchronize()
ws = eng._last_ws
p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
col = {}
enc, olen = TT.tdnn_encode(feats.double(), lens.long(), MINI_JASPER, p64, emulate_storage=True, collect=col)
for t in col.values():
    t.retain_grad()
ref_logits = TT.fc_decode(enc, p64["fc/kernel"], p64["fc/bias"]).transpose(0, 1)
print("== forward (engine vs emulated twin): max-rel / l2-rel")
for li, l in enumerate(eng.layers):
    y = ws.Y[li].float().cpu()
    a = ws.A[li].float().cpu()
    print("%-8s Y %.2e %.2e   A %.2e %.2e" % (l.name, _rel(y, col[l.name + "/conv"]), _rel_l2(y, col[l.name + "/conv"]),
                                              _rel(a, col[l.name + "/out"]), _rel_l2(a, col[l.name + "/out"])))
    mi = ws.mean_invstd[ws.bn_slot[li][0][0]].cpu().reshape(-1)[: 2 * l.c_out].view(2, l.c_out)
    c = col[l.name + "/conv"].detach()
    C = l.c_out
    m_ref = c.mean((0, 1)); v_ref = c.var((0, 1), unbiased=False)
    print("         mean err %.2e  invstd rel err %.2e" % (float((mi[0, :C] - m_ref).abs().max()),
          float(((mi[1, :C] - torch.rsqrt(v_ref + 1e-3)) / torch.rsqrt(v_ref + 1e-3)).abs().max())))
print("logits", _rel(logits.cpu(), ref_logits.detach()), _rel_l2(logits.cpu(), ref_logits.detach()))

g = torch.Generator().manual_seed(9)
R = torch.randn(logits.shape, generator=g)
R = R * TT.sequence_mask(out_lens.cpu().long(), logits.shape[1], R.dtype)
(ref_logits * R.double()).sum().backward()
# step the engine backward manually, layer by layer, dumping dY / dA

from openseq2seq_b200 import _lib as L
ws.dlogits.copy_(R.cuda())
if ws._bwd_plan is None:
    ws._build_backward_plan(32)
plan = ws._bwd_plan[1:]
print("== backward: per launch")
li = len(eng.layers) - 1
for entry in plan:
    fn, args = entry[0], entry[1]
    name = fn.__name__
    L.check(fn(*args), name)
    torch.cuda.synchronize()
    if name == "os2s_fc_bwd":
        dA = ws.dA.view(-1)[: ws.M * eng.H].view(ws.B, ws.T2, eng.H).float().cpu()
        r = col[eng.layers[-1].name + "/out"].grad
        print("fc_bwd dA(last) %.2e %.2e" % (_rel(dA, r), _rel_l2(dA, r)))
    if name == "os2s_bn_bwd":
        l = eng.layers[li]
        dY = ws.dY2[li & 1].view(-1)[: ws.M * l.c_out].view(ws.B, ws.T2, l.c_out).float().cpu()
        r = col[l.name + "/conv"].grad
        print("%-8s bn_bwd dY %.2e %.2e" % (l.name, _rel(dY, r), _rel_l2(dY, r)), end="")
        for nm in (l.name + "/bn/gamma", l.name + "/bn/beta"):
            print("  %s %.2e" % (nm.split("/")[-1], _rel_l2(eng.param_view(nm, eng.grad).cpu(), p64[nm].grad)), end="")
        print()
        li -= 1
print("== final parameter gradient l2-rel errors")
for name, _ in eng.named_parameters():
    print("%-28s %.3e" % (name, _rel_l2(eng.param_view(name, eng.grad), p64[name].grad)))
