"""Diagnostic: per-parameter gradient error of the CUDA MLP vs the oracle (run on a GPU box)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import robust_nerf_b200 as rn
from oracle import nerf_oracle as O
from conftest import load_golden
dev = torch.device("cuda:0")
def T(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def N(t): return t.detach().float().cpu().numpy()
for tag in ("plain", "sharp"):
    g = load_golden(f"nerf_{tag}")
    w = O.make_weights(7, sharpen=(tag == "sharp"))
    net = rn.NeRF().to(dev)
    sd = net.state_dict()
    for k, v in w.items(): sd[k] = T(v)
    net.load_state_dict(sd)
    x, d = T(g["pts"]).requires_grad_(True), T(g["dirs"]).requires_grad_(True)
    rgb, sigma = net(x, d)
    (rgb * T(g["g_rgb"])).sum().add((sigma * T(g["g_sigma"])).sum()).backward()
    _, _, cache = O.nerf_forward(w, g["pts"], g["dirs"], keep_cache=True, emulate_bf16="--emulate" in sys.argv)
    grads, dx, dd = O.nerf_backward(w, cache, g["g_rgb"], g["g_sigma"], need_input_grad=True)
    print(tag, "rgb err", np.abs(N(rgb) - g["rgb"]).max(), "sigma err", np.abs(N(sigma) - g["sigma"]).max(), "sigma max", g["sigma"].max())
    for k, p in net.named_parameters():
        ref = grads[k]; got = N(p.grad)
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-12)
        print(f"  {k:28s} rel {rel:.4f}  |ref| {np.linalg.norm(ref):.3e} |got| {np.linalg.norm(got):.3e}")
    for nm, a, b in (("dx", x.grad, dx), ("dd", d.grad, dd)):
        print(f"  {nm} rel {np.linalg.norm(N(a) - b) / np.linalg.norm(b):.4f}")
    if tag == "plain":
        got = N(net.pts_linears[0].weight.grad); ref = grads["pts_linears.0.weight"]
        colerr = np.linalg.norm(got - ref, axis=0) / (np.linalg.norm(ref, axis=0) + 1e-12)
        print("  W0 per-column rel err:", np.round(colerr, 3))
