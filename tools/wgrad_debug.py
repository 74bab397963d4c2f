"""Debug the tensor-core wgrad kernel on crafted inputs (diagnostic)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from playablevideogeneration_b200 import _lib
from playablevideogeneration_b200._lib import ConvDesc, call
dev = "cuda"

def run(x, g, r, cin_log, nprod):
    n, h, w, cinp = x.shape
    cout = g.shape[3]
    scratch = torch.zeros(cout * r * r * cinp, device=dev)
    dw = torch.zeros(cout, cin_log, r, r, device=dev)
    d = ConvDesc(n, h, w, cinp, cout, r, r, (r - 1) // 2, 0, 0.0, 1, nprod)
    lo_x = torch.zeros_like(x) if nprod == 3 else None
    lo_g = torch.zeros_like(g) if nprod == 3 else None
    call("pvg_conv2d_wgrad_umma", d, cin_log, x.data_ptr(), None if lo_x is None else lo_x.data_ptr(), g.data_ptr(),
         None if lo_g is None else lo_g.data_ptr(), scratch.data_ptr(), dw.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return dw, scratch

def ref(x, g, r):
    # dW[co][ci][r][s] = sum_p g[p][co] x[p+tap][ci]
    xx = x.permute(0, 3, 1, 2); gg = g.permute(0, 3, 1, 2)
    w = torch.zeros(g.shape[3], x.shape[3], r, r, device=dev, requires_grad=True)
    y = torch.nn.functional.conv2d(xx, w, padding=r // 2)
    y.backward(gg)
    return w.grad

for nprod in (1, 3):
    for (n, h, w, cin, cout, r) in [(1, 8, 4, 32, 32, 1), (1, 8, 4, 32, 32, 3), (2, 8, 8, 64, 128, 1), (1, 16, 16, 32, 64, 3)]:
        torch.manual_seed(0)
        tests = {
            "ones": (torch.ones(n, h, w, cin, device=dev), torch.ones(n, h, w, cout, device=dev)),
            "g=co": (torch.ones(n, h, w, cin, device=dev), torch.arange(cout, device=dev, dtype=torch.float32).expand(n, h, w, cout).contiguous()),
            "x=ci": (torch.arange(cin, device=dev, dtype=torch.float32).expand(n, h, w, cin).contiguous(), torch.ones(n, h, w, cout, device=dev)),
            "rand": (torch.randn(n, h, w, cin, device=dev), torch.randn(n, h, w, cout, device=dev)),
        }
        for name, (x, g) in tests.items():
            dw, scratch = run(x, g, r, cin, nprod)
            rf = ref(x, g, r)
            err = float((dw - rf).abs().max()); sc = float(rf.abs().max())
            print(f"nprod={nprod} shape={(n,h,w,cin,cout,r)} {name:5s}: err {err:.3e} / {sc:.3e}  scratch absmax {float(scratch.abs().max()):.3e} "
                  f"nonzero {int((scratch != 0).sum())}/{scratch.numel()}  dw[0,:4,0,0]={dw[0,:4,0,0].tolist()} ref={rf[0,:4,0,0].tolist()} dw[:4,0,0,0]={dw[:4,0,0,0].tolist()}")
