import os, sys, subprocess, textwrap
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRE = """
import sys, torch, traceback
sys.path.insert(0, %r); sys.path.insert(0, %r + '/oracle')
import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
B, N = 64, 256
x = (torch.randn(B, N, 2, generator=g)).to(dev)
ctx = torch.randn(B, 36, generator=g).to(dev)
pk = O.init_stack(g, 2, 36, std=0.1, bias_std=0.05).to(dev).requires_grad_()
lin = torch.nn.Linear(2, 2).to(dev)
def capture(step):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): step()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    pk.grad = None
    for p in lin.parameters(): p.grad = None
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step()
    gr.replay(); torch.cuda.synchronize()
""" % (ROOT, ROOT)
CASES = {
 "torch_only": "def step():\n    (lin(x).sum()).backward()\ncapture(step)",
 "coupling_param": "def step():\n    y, ld = ops.coupling_stack(pk, x, ctx, None, 2, True)\n    (y.sum()+ld.sum()).backward()\ncapture(step)",
 "coupling_after_torch": "def step():\n    y, ld = ops.coupling_stack(pk, lin(x), ctx, None, 2, True)\n    (y.sum()+ld.sum()).backward()\ncapture(step)",
 "cat_param": "ws=[torch.randn(10,device=dev).requires_grad_() for _ in range(3)]\ndef step():\n    (torch.cat(ws)*2).sum().backward()\ncapture(step)",
 "soft": "w=torch.softmax(torch.randn(B,N,device=dev),-1).requires_grad_()\nmk=torch.linspace(0,(N-1)/N,N).to(dev); off=(torch.rand(B)/N).to(dev)\ndef step():\n    p,w2,i,l=ops.soft_resample(x,w,off,mk,0.5,True)\n    (p.sum()+l.sum()).backward()\ncapture(step)",
}
for name, body in CASES.items():
    r = subprocess.run([sys.executable, "-c", PRE + body], capture_output=True, text=True)
    tail = [l for l in r.stderr.strip().splitlines() if "Error" in l or "error" in l][-2:]
    print(name, "OK" if r.returncode == 0 else "FAIL", tail)
