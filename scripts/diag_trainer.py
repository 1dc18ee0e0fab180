import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import robust_nerf_b200 as rn
from oracle import nerf_oracle as O
dev = torch.device("cuda:0")
def load_net(w):
    net = rn.NeRF().to(dev); sd = net.state_dict()
    for k, v in w.items(): sd[k] = torch.from_numpy(v).to(dev)
    net.load_state_dict(sd); return net
def two():
    return load_net(O.make_weights(41, sharpen=True)), load_net(O.make_weights(42, sharpen=True))
rng = np.random.default_rng(0)
data = rn.make_scene(64, 64, 100, seed=3, device=dev)
ds, sampler = rn.create_pixel_dataset(data)
pb = sampler.batch_from_indices(torch.from_numpy(rng.integers(0, ds.n_pixels, 256)).to(dev))
with torch.no_grad(): ro, rd = sampler.get_rays_for_batch(pb, data.poses)
cfg = rn.RenderConfig()
def grads_once(nc, nf, seed):
    torch.manual_seed(seed)
    for p in list(nc.parameters()) + list(nf.parameters()): p.grad = None
    out = rn.render_rays(nc, nf, ro, rd, cfg, is_train=True)
    loss = ((out["rgb_coarse"] - pb.target_rgb) ** 2).mean() + ((out["rgb_fine"] - pb.target_rgb) ** 2).mean()
    loss.backward()
    return loss.item(), torch.cat([p.grad.reshape(-1) for p in list(nc.parameters()) + list(nf.parameters())]).clone()
nc, nf = two()
l1, g1 = grads_once(nc, nf, 5); l2, g2 = grads_once(nc, nf, 5)
print("determinism: loss equal", l1 == l2, "grad max diff", (g1 - g2).abs().max().item(), "gnorm", g1.norm().item())
# one step each way
nc, nf = two(); renderer = rn.NeRFRenderer(nc, nf, cfg)
opt = torch.optim.Adam(renderer.parameters(), lr=5e-4)
torch.manual_seed(100); ma = rn.train_step(renderer, opt, {"rays_o": ro, "rays_d": rd, "target_rgb": pb.target_rgb})
ga = torch.cat([p.grad.reshape(-1) for p in renderer.parameters()]).clone()
mc, mf = two(); tr = rn.Trainer(mc, mf, cfg, lr=5e-4, lr_decay_steps=1e30)
torch.manual_seed(100); lb = tr.step_rays(ro, rd, pb.target_rgb).item()
gb = tr.gflat.clone()
print("loss A", ma["loss"], "loss B", lb)
print("clipped grad diff", (ga - gb).abs().max().item(), "norm A", ga.norm().item(), "norm B", gb.norm().item(), "norms kernel", tr.norms[:2].tolist())
pa = torch.cat([p.detach().reshape(-1) for p in renderer.parameters()]); pbb = tr.flat[:pa.numel()]
d = (pa - pbb).abs(); i = d.argmax().item()
print("param diff max", d.max().item(), "at", i, "ga", ga[i].item(), "gb", gb[i].item(), "m", tr.exp_avg[i].item(), "v", tr.exp_avg_sq[i].item())
print("frac > 1e-6:", (d > 1e-6).float().mean().item())
