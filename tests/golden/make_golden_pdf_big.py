"""Reference `sample_pdf` indices and cdf at the benchmark's batch size.

Runs the UNMODIFIED reference (`/root/reference/noisy_src/rays.py:213-279`) on the CPU of the authoring container
and records what its own `torch.searchsorted(cdf, u, right=True)` call received and returned, by wrapping
`torch.searchsorted` for the duration of the call (the reference does not return `inds` / `cdf`).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_pdf_big.py

Inputs are NOT stored: `pdf_big_inputs(case)` below rebuilds them from a seeded numpy generator, and the tests
import it.  Stored per case: `inds` for every ray (uint8), `samples` (fp16-free: fp32) and `cdf` for every 8th ray.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"c64_f128": (4096, 64, 128, 11), "c128_f256": (2048, 128, 256, 12)}
CDF_STRIDE = 8


def pdf_big_inputs(case):
    """z [B,Nc] (stratified, sorted), weights [B,Nc] (half of the rays: U^4 spikes, the other half: compositing
    weights of a random density field, the shape of noisy_src/test_baseline.py:112-116), u [B,Nf] ~ U[0,1)."""
    B, Nc, Nf, seed = CASES[case]
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 1.0, Nc, dtype=np.float32)
    zb = (np.float32(2.0) * (np.float32(1.0) - t) + np.float32(6.0) * t).astype(np.float32)
    mids = (np.float32(0.5) * (zb[1:] + zb[:-1])).astype(np.float32)
    lower = np.concatenate([zb[:1], mids]); upper = np.concatenate([mids, zb[-1:]])
    z = (lower + (upper - lower) * rng.uniform(0, 1, (B, Nc)).astype(np.float32)).astype(np.float32)
    w = (rng.uniform(0, 1, (B, Nc)) ** 4).astype(np.float32)
    sigma = (10.0 * rng.uniform(0, 1, (B // 2, Nc)) * (rng.uniform(0, 1, (B // 2, Nc)) < 0.15)).astype(np.float32)
    d = np.concatenate([z[B // 2:, 1:] - z[B // 2:, :-1], np.full((B // 2, 1), 1e10, np.float32)], -1)
    alpha = 1.0 - np.exp(-sigma * d)
    T = np.cumprod(np.concatenate([np.ones((B // 2, 1)), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    w[B // 2:] = (alpha * T).astype(np.float32)
    w[0] = 0.0
    w[1] = 0.0; w[1, Nc // 3] = 5.0
    u = rng.uniform(0, 1, (B, Nf)).astype(np.float32)
    return z, w, u


def main():
    import torch
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from noisy_src import rays as R
    torch.set_num_threads(8)
    out = {}
    for case, (B, Nc, Nf, _) in CASES.items():
        z, w, u = pdf_big_inputs(case)
        zt, wt, ut = map(torch.from_numpy, (z, w, u))
        mids = 0.5 * (zt[..., 1:] + zt[..., :-1])
        rec = {}
        orig_ss, orig_rand = torch.searchsorted, torch.rand

        def ss(cdf, uu, **kw):
            r = orig_ss(cdf, uu, **kw)
            rec.update(cdf=cdf.clone(), u=uu.clone(), inds=r.clone())
            return r
        torch.searchsorted, torch.rand = ss, (lambda *a, **k: ut.clone())
        try:
            samples = R.sample_pdf(mids, wt[..., 1:-1], Nf, det=False)
        finally:
            torch.searchsorted, torch.rand = orig_ss, orig_rand
        assert torch.equal(rec["u"], ut) and int(rec["inds"].max()) < 256
        out[f"{case}_inds"] = rec["inds"].numpy().astype(np.uint8)
        out[f"{case}_samples"] = samples.numpy()[::CDF_STRIDE].copy()
        out[f"{case}_cdf"] = rec["cdf"].numpy()[::CDF_STRIDE].copy()
        print(case, "inds", rec["inds"].shape, "max", int(rec["inds"].max()), "cdf[-1] range",
              float(rec["cdf"][:, -1].min()), float(rec["cdf"][:, -1].max()))
    path = os.path.join(HERE, "sample_pdf_big.npz")
    np.savez_compressed(path, **out)
    print(f"sample_pdf_big.npz: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
