"""CPU-only tests: the C-ABI library builds, loads and exports every symbol the header declares
(no compute calls without a GPU); host-side logic (configs, state_dict contract, synthetic scene,
no-fallback behaviour); and the N>1 data-parallel / tile-sharding logic under gloo, world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import nerf_oracle as O


@pytest.fixture(scope="module")
def rn():
    import __graft_entry__ as ge
    ge.build()
    import robust_nerf_b200 as m
    return m


def test_library_exports_every_header_symbol(rn):
    from robust_nerf_b200 import _lib
    header = open(os.path.join(ROOT, "include", "rnerf_b200.h")).read()
    declared = set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.lib()                       # dlopen + resolve all symbols
    assert lib.rn_version() == 100
    assert b"invalid" in lib.rn_status_string(1)
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rn_[a-z0-9_]+)", nm))
    assert declared <= exported
    assert lib.rn_mlp_packed_weight_bytes() > 595844 * 2
    assert lib.rn_mlp_workspace_bytes(4096 * 64, 1) > lib.rn_mlp_workspace_bytes(4096 * 64, 0)


def test_sass_uses_blackwell_tensor_path(rn):
    """tcgen05.mma / tcgen05.ld / TMA must be in the shipped SASS (UTCHMMA / LDTM / UTMALDG / UTMASTG)."""
    from robust_nerf_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    assert "UTCHMMA.2CTA" in sass          # the chain kernels issue cta_group::2 MMAs (M = 256 per CTA pair)
    # packed two-column epilogue arithmetic: bias add, convert+ReLU, mask flags
    assert "FADD2" in sass and "F2FP.RELU" in sass and "VIMNMX.U16x2" in sass
    assert "HMMA.16816" not in sass        # no legacy mma.sync path


def test_weight_gradient_stream_plan(rn):
    """Host logic of csrc/wgrad_stream.cu: the ten weight-gradient GEMMs of a backward pass share the stream's CTA pairs
    so that the slowest pair (bytes per 64-point chunk / splits) is as fast as possible; every GEMM gets at least one
    pair and all pairs are used.  No device work."""
    import ctypes
    from robust_nerf_b200 import _lib
    lib = _lib.lib()
    splits, n = (ctypes.c_int * 11)(), ctypes.c_int(0)
    assert lib.rn_debug_stream_plan(88, splits, ctypes.byref(n)) == 0
    assert n.value == 10                                   # dir, feature, L7 .. L1, L0 (sigma_linear is heads_bwd's)
    s = list(splits)[:10]
    assert sum(s) == 44 and min(s) >= 4, s
    assert s[0] >= 5 and s[4] >= 5, s                      # the two 320-wide GEMMs (dir_linear, layer 5) pull the most bytes per chunk
    assert lib.rn_debug_stream_plan(20, splits, ctypes.byref(n)) == 0 and list(splits)[:10] == [1] * 10
    assert lib.rn_debug_stream_plan(18, splits, ctypes.byref(n)) != 0          # fewer pairs than GEMMs


def test_configs_match_reference_defaults(rn):
    for ours, ref in ((rn.ModelConfig(), O.ModelConfig()), (rn.RenderConfig(), O.RenderConfig())):
        assert vars(ours) == vars(ref)
    cfg = rn.NeRFConfig()
    assert cfg.train.lr == 5e-4 and cfg.train.seed == 42 and cfg.data.batch_size == 1024 and cfg.pose_opt is None


def test_state_dict_contract_and_init_parity(rn):
    torch.manual_seed(42)
    net = rn.NeRF()
    sd = net.state_dict()
    shapes = O.param_shapes(O.ModelConfig())
    assert list(sd.keys()) == ["pos_encoder.freq_bands", "dir_encoder.freq_bands"] + O.param_names(O.ModelConfig())
    assert all(tuple(sd[k].shape) == s and sd[k].dtype == torch.float32 for k, s in shapes.items())
    assert torch.equal(sd["pos_encoder.freq_bands"], 2.0 ** torch.arange(10.0))
    assert sum(p.numel() for p in net.parameters()) == 595844
    # parameters stay ordinary fp32 leaf nn.Parameters (Adam / clip_grad_norm_ operate on them)
    assert all(isinstance(p, torch.nn.Parameter) and p.is_leaf for p in net.parameters())
    # same construction order as the reference => same default init under the same seed
    torch.manual_seed(42)
    lin = torch.nn.Linear(63, 256)
    assert torch.equal(lin.weight, sd["pts_linears.0.weight"])
    cam = rn.CameraPoseParameters(torch.eye(4).repeat(5, 1, 1))
    assert list(cam.state_dict().keys()) == ["rotation_deltas", "translation_deltas", "initial_poses"]  # as the reference: parameters, then buffers
    assert [n for n, _ in cam.named_parameters()] == ["rotation_deltas", "translation_deltas"]
    assert cam.n_poses == 5 and cam.learn_rotation and cam.learn_translation
    cam2 = rn.CameraPoseParameters(torch.eye(4).repeat(5, 1, 1), learn_rotation=False)
    assert [n for n, _ in cam2.named_parameters()] == ["translation_deltas"]


def test_no_cpu_fallback(rn):
    net = rn.NeRF()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        net(torch.randn(8, 3), torch.randn(8, 3))
    with pytest.raises(RuntimeError):
        rn.sample_along_rays(torch.randn(4, 3), torch.randn(4, 3), 2.0, 6.0, 8, perturb=False)
    with pytest.raises(NotImplementedError):
        rn.NeRF(rn.ModelConfig(num_hidden_layers=4))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or the reference)."""
    pkg = os.path.join(ROOT, "robust-nerf_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src and "/root/reference" not in src and "noisy_src" not in src.replace(
                "noisy_src/", "").replace("noisy_src.", ""), f


def test_synthetic_scene_helpers(rn):
    poses = rn.lego_poses()
    assert poses.shape == (100, 4, 4)
    np.testing.assert_allclose(torch.norm(poses[:, :3, 3], dim=-1).numpy(), 4.0311, atol=1e-3)
    assert torch.equal(poses, torch.from_numpy(load_golden("lego_poses")["ground_truth_poses"]))
    hp = rn.hemisphere_poses(7, seed=1)
    R = hp[:, :3, :3]
    np.testing.assert_allclose((R @ R.transpose(-1, -2)).numpy(), np.broadcast_to(np.eye(3), (7, 3, 3)), atol=1e-6)
    assert (hp[:, 2, 3] > 0).all()
    from robust_nerf_b200.synthetic import focal_from_fov
    assert abs(focal_from_fov(800) - 1111.111) < 1e-2 and abs(focal_from_fov(100) - 138.889) < 1e-2
    # the "5 deg / 5 %" noisy initialisation equals the one the reference generated (golden pose.npz): host-side draws
    # in the reference's order; the arithmetic itself is CUDA-only (tests/test_gpu_parity.py) and checked here through
    # the numpy restatement
    from oracle import nerf_oracle as O
    ga, gx, gt_ = rn.draw_pose_noise(poses, rn.NoiseConfig(5.0, 0.0, 5.0, seed=42))
    noisy, _ = O.add_noise_to_poses(poses.numpy(), ga.numpy(), gx.numpy(), gt_.numpy(), 5.0, 0.0, 5.0)
    np.testing.assert_allclose(noisy, load_golden("pose")["init"], atol=1e-6)


def test_tile_and_batch_sharding(rn):
    from robust_nerf_b200.parallel import shard_range, tiles_for_rank
    assert [shard_range(4096, r, 4) for r in range(4)] == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)]
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    tiles_per_view = 20        # 640,000 rays / 32,768-ray tiles
    for world in (1, 2, 4, 8):
        owned = [(v, k) for r in range(world) for v in range(6) for k in tiles_for_rank(v, tiles_per_view, r, world)]
        assert sorted(owned) == [(v, k) for v in range(6) for k in range(tiles_per_view)]     # exact cover
        per_rank = [sum(len(tiles_for_rank(v, tiles_per_view, r, world)) for v in range(6)) for r in range(world)]
        assert max(per_rank) - min(per_rank) <= 1


_DP_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT={port!r}, RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
import importlib.util
spec = importlib.util.spec_from_file_location("par", os.path.join({root!r}, "robust-nerf_b200", "parallel.py"))
par = importlib.util.module_from_spec(spec); spec.loader.exec_module(par)
torch.manual_seed(0)
w = torch.randn(50, requires_grad=True)          # replicated "parameters"
x = torch.randn(64, 50); y = torch.randn(64)     # global batch, same on every rank (one seed, then slice)
a, b = par.shard_range(64, rank, world)
loss = ((x[a:b] @ w - y[a:b]) ** 2).mean()       # mean over the LOCAL shard
loss.backward()
flat = w.grad.clone()                            # flat gradient buffer
par.allreduce_mean_(flat, world)
ref = torch.autograd.grad(((x @ w.detach().requires_grad_(True) - y) ** 2).mean(), [])  if False else None
w2 = w.detach().clone().requires_grad_(True)
((x @ w2 - y) ** 2).mean().backward()
assert torch.allclose(flat, w2.grad, atol=1e-6), (flat - w2.grad).abs().max()
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(g, gathered[0]) for g in gathered)   # every rank steps with the same gradient
dist.destroy_process_group()
print("ok", rank)
"""


def test_data_parallel_gradient_average_gloo_world2(tmp_path):
    """world_size-2 gloo run of the N>1 path: shard the batch, ONE all-reduce of the flat gradient
    buffer, 1/world scaling == single-process global-batch gradient; all ranks end identical."""
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER.format(root=ROOT, port=str(29500 + os.getpid() % 2000)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


NOISE_CASES = ("rot5_pct5", "rot2_abs", "pct3", "rot1p5", "clean")


def test_pose_noise_oracle_against_reference(rn):
    """add_noise_to_poses (noisy_src/noise.py:71-234) restated in the oracle and fed with the
    package's host-side draw routine (the reference's generator call order), against noisy poses and noise_info
    produced by the unmodified reference (tests/golden/make_golden_noise.py); compute_pose_error (noise.py:237-268)
    against the reference's per-pose errors."""
    import torch
    g = load_golden("noise")
    poses = g["poses"]
    for tag in NOISE_CASES:
        rot, tabs, pct, seed = g[f"{tag}_cfg"]
        cfg = rn.NoiseConfig(float(rot), float(tabs), float(pct), seed=int(seed))
        ga, gx, gt_ = rn.draw_pose_noise(torch.from_numpy(poses), cfg)
        noisy, info = O.add_noise_to_poses(poses, None if ga is None else ga.numpy(), None if gx is None else gx.numpy(),
                                           None if gt_ is None else gt_.numpy(), float(rot), float(tabs), float(pct))
        close(noisy, g[f"{tag}_noisy"], rtol=0, atol=1e-6)
        # angle = acos((trace - 1) / 2): one ulp of the trace is ~1e-4 degrees at 1 degree, more below
        close(info[:, 0], g[f"{tag}_info"][:, 0], rtol=1e-4, atol=5e-3)
        close(info[:, 1], g[f"{tag}_info"][:, 1], rtol=1e-5, atol=1e-6)
        err = np.array([[e["rotation_error_deg"], e["translation_error"]]
                        for e in (O.compute_pose_error(poses[i], g[f"{tag}_noisy"][i]) for i in range(len(poses)))])
        # acos near 1 turns one ulp of the trace into ~0.03 degrees: absolute tolerance for the rotation error
        close(err[:, 0], g[f"{tag}_err"][:, 0], rtol=1e-4, atol=0.06)
        close(err[:, 1], g[f"{tag}_err"][:, 1], rtol=1e-5, atol=1e-6)
    # the pose at the origin got no translation noise under percentage noise (std = 0: no draw, noise.py:184)
    assert np.array_equal(g["rot5_pct5_noisy"][7, :3, 3], np.zeros(3, np.float32))
    assert str(rn.NoiseConfig(5.0, 0.0, 5.0)) == "rot5.0deg_trans5.0pct" and str(rn.NoiseConfig()) == "clean"
    assert rn.NoiseConfig(0.0, 0.1).has_noise and not rn.NoiseConfig().has_noise
    assert rn.NoiseConfig(0, 0.2, 0).get_translation_std(4.0) == 0.2 and rn.NoiseConfig(0, 0.2, 5.0).get_translation_std(4.0) == 0.2


def test_reference_arm_runs_and_reports_all_host_threads():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) executes here: it times the reference's fp32
    PyTorch training step on the host and prints the contract line.  Run under OMP_NUM_THREADS=1, as torchrun sets it for
    N > 1: the arm must still use every core it is allowed on.  RN_BENCH_CPU_RAYS shrinks the
    sample so the CPU suite stays short; RN_BENCH_FORCE_PORT selects the pinned torch port as on the GPU box."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1", RN_BENCH_FORCE_PORT="1", RN_BENCH_CPU_RAYS="64", RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--gpus", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0 and line["vs_baseline"] is None
    c = line["cpu_baseline"]
    try:
        allowed = len(os.sched_getaffinity(0))
    except AttributeError:
        allowed = os.cpu_count()
    assert c["kind"] == "port" and c["cores"] == allowed and c["value"] == line["value"] and c["render"]["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 exit 0 without work
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=dict(env, RANK="1"), timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
