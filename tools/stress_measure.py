"""Run-to-run determinism stress of the measurement kernels: nine shapes (all three likelihoods, fused and not, few and many
trajectories per persistent CTA) replayed 400 times in random order in one process, every output and gradient compared bitwise with
the first evaluation of its case.  B200: 0 mismatches."""
import sys, torch, random
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
cases = [("gaussian", 3, 200, True), ("cos", 3, 200, True), ("CRNVP", 3, 200, True), ("gaussian", 16, 1024, True), ("CRNVP", 8, 1024, True),
         ("gaussian", 700, 200, True), ("cos", 650, 130, False), ("CRNVP", 600, 129, True), ("CRNVP", 300, 1024, True)]
data = {}
for ci, (mode, B, N, fused) in enumerate(cases):
    g = torch.Generator().manual_seed(100 + ci)
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))]).cuda()
    cnf = O.init_stack(g, 32, 32, std=0.1, bias_std=0.05).cuda() if mode == "CRNVP" else None
    t = [pe, cnf, torch.randn(B, 32, generator=g).cuda(), (torch.randn(B, N, 2, generator=g) * 3).cuda(),
         torch.log_softmax(torch.randn(B, N, generator=g), -1).cuda(), torch.randn(B, N, generator=g).cuda(), torch.randn(B, N, generator=g).cuda()]
    gs = [torch.randn(B, N, generator=g).cuda(), torch.randn(B, N, generator=g).cuda(), torch.randn(B, generator=g).cuda()]
    data[ci] = (t, gs)
def run(ci):
    mode, B, N, fused = cases[ci]
    t, gs = data[ci]
    tt = [v.clone().requires_grad_() if v is not None else None for v in t]
    p0, p1 = {"gaussian": (1.0, 10.0), "cos": (0.0, 1.0), "CRNVP": (0.0, 2.5)}[mode]
    if fused:
        lki, logw, probs, rs, ess = ops.measure_update(tt[0], tt[1], tt[2], tt[3], tt[4], tt[5], tt[6], mode, p0=p0, p1=p1)
        ((lki * gs[0]).sum() + (probs * gs[1]).sum() * 50 + (rs * gs[2]).sum() * 0.01).backward()
    else:
        lki = ops.measure(tt[0], tt[1], tt[2], tt[3], mode, p0=p0, p1=p1)
        (lki * gs[0]).sum().backward()
    return [lki.detach().clone()] + [v.grad.clone() for v in tt if v is not None and v.grad is not None]
ref = {ci: run(ci) for ci in range(len(cases))}
random.seed(0)
bad = 0
for it in range(400):
    ci = random.randrange(len(cases))
    cur = run(ci)
    for k, (a, b) in enumerate(zip(cur, ref[ci])):
        d = (a - b).abs().max().item()
        if d != 0:
            bad += 1
            print("MISMATCH iter", it, cases[ci], "tensor", k, "max diff", d)
print("stress done, mismatches:", bad)
