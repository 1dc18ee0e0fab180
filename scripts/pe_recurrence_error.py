#!/usr/bin/env python
"""Error of the angle-doubling positional encoding (csrc/pe.cuh) against float64 sin/cos, per frequency, for different
restart schedules (accurate sincosf at the listed k, two-term doubling in between).  CPU, numpy fp32.
    python scripts/pe_recurrence_error.py > profiles/r02_pe_recurrence.md"""
import numpy as np

F = np.float32
rng = np.random.default_rng(0)
x = rng.uniform(-6, 6, 200000).astype(F)


def bf16(v):
    u = np.ascontiguousarray(v, dtype=F).view(np.uint32)
    r = ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).astype(np.uint32)
    return r.view(F)


def run(restarts, L=10):
    out, s, c = [], None, None
    for k in range(L):
        if k in restarts:
            a = (F(2.0 ** k) * x).astype(F)
            s, c = np.sin(a.astype(np.float64)).astype(F), np.cos(a.astype(np.float64)).astype(F)
        else:
            s2 = (s + s).astype(F)
            cn = (F(1.0) - (s2 * s).astype(F)).astype(F)
            s, c = (s2 * c).astype(F), cn
        ts, tc = np.sin((2.0 ** k) * x.astype(np.float64)), np.cos((2.0 ** k) * x.astype(np.float64))
        err = max(np.abs(s - ts).max(), np.abs(c - tc).max())
        flip = ((bf16(s) != bf16(ts.astype(F))).mean() + (bf16(c) != bf16(tc.astype(F))).mean()) / 2
        out.append((err, flip))
    return out


def run_turns(L=10, mufu_err=2.0 ** -21.41):
    """csrc/pe.cuh as built: two-float turn reduction, then sin / cos of the reduced angle with the documented absolute
    error bound of sin.approx / cos.approx on [-pi, pi] added as uniform noise."""
    c = 1.0 / (2.0 * np.pi)
    c_hi = F(c); c_lo = F(c - float(c_hi))
    t_hi = (x * c_hi).astype(F)
    resid = (x.astype(np.float64) * float(c_hi) - t_hi.astype(np.float64)).astype(F)
    t_lo = (x.astype(np.float64) * float(c_lo) + resid.astype(np.float64)).astype(F)
    out = []
    for k in range(L):
        f = F(2.0 ** k)
        y = (t_hi * f).astype(F)
        r = (t_lo.astype(np.float64) * float(f) + (y - np.rint(y)).astype(F).astype(np.float64)).astype(F)
        ang = (r * F(6.2831855)).astype(F).astype(np.float64)
        s = (np.sin(ang) + rng.uniform(-mufu_err, mufu_err, ang.shape)).astype(F)
        c_ = (np.cos(ang) + rng.uniform(-mufu_err, mufu_err, ang.shape)).astype(F)
        ts, tc = np.sin((2.0 ** k) * x.astype(np.float64)), np.cos((2.0 ** k) * x.astype(np.float64))
        err = max(np.abs(s - ts).max(), np.abs(c_ - tc).max())
        flip = ((bf16(s) != bf16(ts.astype(F))).mean() + (bf16(c_) != bf16(tc.astype(F))).mean()) / 2
        out.append((err, flip))
    return out


print("# Fast positional encoding for the bf16 operand: error per frequency (200,000 points in [-6, 6], fp32 vs float64)\n")
print("`max abs err` against float64 sin/cos of 2^k x; `bf16 flips` = fraction of features whose bf16 rounding differs from")
print("the rounding of the exact value.  BUILT (csrc/pe.cuh): the two-float turn reduction of the first table; the")
print("angle-doubling schedules below were the alternatives considered (same accuracy class, more instructions).\n")
for name, rs in (("two-float turn reduction + sin.approx / cos.approx (BUILT)", None), ("accurate sincosf at every k (round 1)", set(range(10))),
                 ("angle doubling, restarts {0, 3, 6, 9}", {0, 3, 6, 9}), ("angle doubling, restarts {0, 5}", {0, 5}),
                 ("angle doubling, restart {0} only", {0})):
    r = run_turns() if rs is None else run(rs)
    print(f"## {name}\n")
    print("| k | " + " | ".join(str(k) for k in range(10)) + " |")
    print("|---|" + "---|" * 10)
    print("| max abs err | " + " | ".join(f"{e:.1e}" for e, _ in r) + " |")
    print("| bf16 flips | " + " | ".join(f"{f:.1e}" for _, f in r) + " |\n")
