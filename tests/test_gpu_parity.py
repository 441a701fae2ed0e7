"""GPU parity tests: the sm_100a kernels (called through the C ABI) against the CPU oracle on
identical seeded inputs, against the committed golden vectors, and through size-independent
properties at full size.

Tolerances (fp32 path; the calibration is the fp32-oracle vs fp64-oracle gap printed by
tests/golden/make_golden.py, <= ~6e-6 relative):
  * per-Gaussian streaming kernels (project, SH): 1e-5 relative to the tensor's max |value|
  * image: 2e-4 absolute on [0,1] colours (ex2.approx + reordered fp32 sums)
  * gradients: 1e-3 relative to the tensor's max |value| (float atomics reorder the sums)
Integer outputs (radii, tile counts) must match exactly except where a ceil/floor sits within
one fp32 ulp of an integer; at most 0.1% of entries may differ, by at most 1."""
import os

import numpy as np
import pytest
import torch

import oracle
from tinysplat_b200 import synthetic

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "config1.npz")
DEV = "cuda:0"
PARAMS = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]

TOL_STREAM = 1e-5
TOL_IMG = 2e-4
TOL_GRAD = 1e-3


@pytest.fixture(scope="module", autouse=True)
def _need_cuda(lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(params=[0, 1], ids=["bwd-warp", "bwd-group"])
def blend_mode(request, lib):
    """Runs a test once per blend-backward kernel (include/tinysplat_b200.h, ts_set_blend_mode):
    0 = first generation (one warp per sub-block), 1 = grouped (default)."""
    assert lib.ts_set_blend_mode(request.param) == 0
    yield request.param
    lib.ts_set_blend_mode(-1)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def proj_inputs(sc, cam, W, H, dtype=torch.float32, device="cpu"):
    V = cam.view_matrix.to(dtype).to(device)
    P = cam.proj_matrix.to(dtype).to(device)
    q = sc["quats"].to(dtype).to(device)
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    return [sc["means"].to(dtype).to(device), torch.exp(sc["scales"].to(dtype)).to(device), 1.0,
            q / q.norm(dim=-1, keepdim=True), V[:3], P @ V, cam.f_x, cam.f_y, W / 2, H / 2, H, W, tb]


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,W,H", [(256, 128, 128), (5000, 640, 360), (1, 16, 16), (777, 37, 21)])
def test_project_forward_and_backward(N, W, H):
    import gsplat
    cam = synthetic.make_camera(W, H, yaw_deg=7.0, shift=(0.3, -0.2, 0.1))
    sc = synthetic.make_scene(N, W, H, seed=N)
    sc["means"][::17, 2] = -1.0          # some behind the camera
    sc["means"][::23, 0] *= 4.0          # some far outside the frustum (exercises the fov clamp)
    ref_in = proj_inputs(sc, cam, W, H, torch.float64)
    for i in (0, 1, 3):
        ref_in[i] = ref_in[i].clone().requires_grad_(True)
    r_xys, r_dep, r_rad, r_con, r_nt, r_cov = oracle.project_gaussians(*ref_in)
    gpu_in = proj_inputs(sc, cam, W, H, torch.float32, DEV)
    for i in (0, 1, 3):
        gpu_in[i] = gpu_in[i].clone().requires_grad_(True)
    xys, dep, rad, con, nt, cov = gsplat.project_gaussians(*gpu_in)
    assert rad.dtype == torch.int32 and nt.dtype == torch.int32
    assert xys.shape == (N, 2) and con.shape == (N, 3) and cov.shape == (N, 6)
    bad = (rad.cpu() != r_rad)
    assert bad.float().mean().item() <= 1e-3 and (rad.cpu() - r_rad).abs().max().item() <= 1
    same = ~bad & ((rad.cpu() > 0) == (r_rad > 0))
    assert (nt.cpu()[same] != r_nt[same]).float().mean().item() <= 1e-3
    for got, want in ((xys, r_xys), (dep, r_dep), (con, r_con), (cov, r_cov)):
        g, w = got.cpu()[same], want[same]
        assert rel_err(g, w) < TOL_STREAM
    # zeros for culled Gaussians
    culled = rad == 0
    assert xys[culled].abs().sum() == 0 and con[culled].abs().sum() == 0 and dep[culled].abs().sum() == 0
    # backward with random cotangents on xys, depths AND conics
    g = torch.Generator().manual_seed(0)
    c_xy = torch.randn(N, 2, generator=g, dtype=torch.float64)
    c_dep = torch.randn(N, generator=g, dtype=torch.float64)
    c_con = torch.randn(N, 3, generator=g, dtype=torch.float64)
    keep = same.double()
    ((r_xys * c_xy * keep[:, None]).sum() + (r_dep * c_dep * keep).sum() + (r_con * c_con * keep[:, None]).sum()).backward()
    kd = keep.float().to(DEV)
    ((xys * c_xy.float().to(DEV) * kd[:, None]).sum() + (dep * c_dep.float().to(DEV) * kd).sum()
     + (con * c_con.float().to(DEV) * kd[:, None]).sum()).backward()
    for i in (0, 1, 3):
        assert torch.isfinite(gpu_in[i].grad).all()
        assert rel_err(gpu_in[i].grad, ref_in[i].grad) < 1e-4, f"input {i}"


@pytest.mark.parametrize("deg,K", [(0, 1), (1, 4), (2, 9), (3, 16), (4, 25), (1, 16), (0, 16), (2, 16)])
def test_spherical_harmonics(deg, K):
    import gsplat
    from tinysplat_b200 import spherical_harmonics_split
    N = 1000 + deg
    g = torch.Generator().manual_seed(deg * 100 + K)
    dirs = torch.randn(N, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    co = torch.randn(N, K, 3, generator=g)
    cot = torch.randn(N, 3, generator=g)
    r_co = co.double().requires_grad_(True)
    want = oracle.spherical_harmonics(deg, dirs.double(), r_co)
    (want * cot.double()).sum().backward()
    g_co = co.to(DEV).requires_grad_(True)
    got = gsplat.sh.spherical_harmonics(deg, dirs.to(DEV), g_co)
    (got * cot.to(DEV)).sum().backward()
    assert rel_err(got, want) < TOL_STREAM
    assert rel_err(g_co.grad, r_co.grad) < TOL_STREAM
    if deg < gsplat.sh.deg_from_sh(K):
        assert g_co.grad[:, (deg + 1) ** 2:, :].abs().max() == 0
    if K > 1:   # (dc, rest) split entry point gives the same numbers
        dc = co[:, 0, :].contiguous().to(DEV).requires_grad_(True)
        rest = co[:, 1:, :].contiguous().to(DEV).requires_grad_(True)
        got2 = spherical_harmonics_split(deg, dirs.to(DEV), dc, rest)
        (got2 * cot.to(DEV)).sum().backward()
        assert torch.equal(got2, got)
        assert torch.equal(dc.grad, g_co.grad[:, 0, :]) and torch.equal(rest.grad, g_co.grad[:, 1:, :])


def _raster_case(N, W, H, seed, CH=3):
    """Projected inputs (from the fp32 oracle, so both sides see identical fp32 values)."""
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(N, W, H, seed=seed)
    xys, dep, rad, con, nt, _ = oracle.project_gaussians(*proj_inputs(sc, cam, W, H, torch.float32))
    g = torch.Generator().manual_seed(seed + 1)
    colors = torch.rand(N, CH, generator=g)
    opac = torch.sigmoid(sc["opacities"])
    bg = torch.rand(CH, generator=g)
    return xys, dep, rad, con, nt, colors, opac, bg


@pytest.mark.parametrize("N,W,H,CH", [(256, 128, 128, 3), (3000, 320, 200, 3), (400, 50, 35, 4),
                                      (300, 64, 64, 1), (1500, 96, 96, 2)])
def test_rasterize_forward_and_backward(N, W, H, CH, blend_mode):
    import gsplat
    xys, dep, rad, con, nt, colors, opac, bg = _raster_case(N, W, H, seed=N + CH, CH=CH)
    leaf = lambda t, dt, dev: t.to(dt).to(dev).clone().requires_grad_(True)
    r = [leaf(xys, torch.float64, "cpu"), leaf(con, torch.float64, "cpu"),
         leaf(colors, torch.float64, "cpu"), leaf(opac, torch.float64, "cpu")]
    want, want_a, aux = oracle.rasterize_gaussians(r[0], dep.double(), rad, r[1], nt, r[2], r[3], H, W,
                                                   bg.double(), return_aux=True)
    c = [leaf(xys, torch.float32, DEV), leaf(con, torch.float32, DEV),
         leaf(colors, torch.float32, DEV), leaf(opac, torch.float32, DEV)]
    got, got_a = gsplat.rasterize_gaussians(c[0], dep.to(DEV), rad.to(DEV), c[1], nt.to(DEV), c[2], c[3],
                                            H, W, bg.to(DEV))
    assert got.shape == (H, W, CH) and got_a.shape == (H, W)
    assert (got.cpu().double() - want).abs().max().item() < TOL_IMG
    assert (got_a.cpu().double() - want_a).abs().max().item() < TOL_IMG
    g = torch.Generator().manual_seed(7)
    w_img = torch.rand(H, W, CH, generator=g)
    w_a = torch.rand(H, W, generator=g)
    ((want * w_img.double()).sum() + (want_a * w_a.double()).sum()).backward()
    ((got * w_img.to(DEV)).sum() + (got_a * w_a.to(DEV)).sum()).backward()
    for name, a, b in zip(("xys", "conics", "colors", "opacity"), c, r):
        assert a.grad.shape == a.shape
        assert torch.isfinite(a.grad).all()
        assert rel_err(a.grad, b.grad) < TOL_GRAD, name


def test_footprint_culling_is_result_invariant(blend_mode):
    """cull_mode=1 (opacity-aware footprint culling at tile and sub-tile level) must not change
    the image or the gradients relative to the plain 3-sigma-bbox algorithm (cull_mode=0)."""
    import gsplat
    from tinysplat_b200 import rasterize as rz
    N, W, H = 4000, 256, 192
    xys, dep, rad, con, nt, colors, opac, bg = _raster_case(N, W, H, seed=5)
    outs = []
    for mode in (0, 1):
        leafs = [t.to(DEV).clone().requires_grad_(True) for t in (xys, con, colors, opac)]
        rz.clear_bin_cache()
        img, a = gsplat.rasterize_gaussians(leafs[0], dep.to(DEV), rad.to(DEV), leafs[1], nt.to(DEV), leafs[2],
                                            leafs[3], H, W, bg.to(DEV), cull_mode=mode)
        M = rz.last_stats["num_intersects"]
        (img.sum() + a.sum()).backward()
        outs.append((img, a, [l.grad for l in leafs], M))
    assert outs[1][3] < outs[0][3], "culling should drop intersections"
    assert outs[0][3] == int(nt.sum()), "cull_mode=0 must emit exactly sum(num_tiles_hit) intersections"
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    for g0, g1 in zip(outs[0][2], outs[1][2]):
        assert rel_err(g1, g0) < 1e-5   # identical terms; only the float-atomic order differs


def test_sorted_lists_match_oracle_order():
    """K3 alone: with cull_mode=0 every tile's id list must equal the oracle's (tile, depth, id)
    order exactly — integer/index work is bit-exact."""
    from tinysplat_b200 import rasterize as rz
    N, W, H = 3000, 200, 120
    xys, dep, rad, con, nt, colors, opac, bg = _raster_case(N, W, H, seed=21)
    dep[::7] = dep[0]      # depth ties: order must fall back to the gaussian id
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    tile, gid = oracle.gsplat_oracle.bin_and_sort(xys, dep, rad, tb)
    edges = torch.searchsorted(tile, torch.arange(tb[0] * tb[1] + 1))
    rz.clear_bin_cache()
    recs, bins = rz.pack_and_bin(xys.to(DEV), dep.to(DEV), rad.to(DEV), con.to(DEV),
                                 opac.reshape(-1).to(DEV), colors.to(DEV), H, W, cull_mode=0, reuse=False)
    assert bins.num_intersects == gid.numel()
    assert torch.equal(bins.tile_offsets.cpu().long(), edges)
    assert torch.equal(bins.ids_sorted.cpu().long()[:gid.numel()], gid)


def test_sort_every_size_class():
    """Per-tile lists of chosen lengths hit every sort path (warp/register sort with 2, 4, 8, 16
    keys per lane; CTA shared-memory classes; lengths at the class edges), with depth ties."""
    from tinysplat_b200 import rasterize as rz
    sizes = [0, 1, 2, 3, 31, 32, 33, 64, 65, 127, 128, 129, 255, 256, 257, 300, 511, 512, 513, 700,
             1024, 2047, 2048, 2049, 3000]
    W, H = 16 * len(sizes), 16
    g = torch.Generator().manual_seed(11)
    xs, tiles = [], []
    for t, n in enumerate(sizes):
        xs.append(torch.rand(n, 2, generator=g) * 10 + 3 + torch.tensor([16.0 * t, 0.0]))
        tiles.append(torch.full((n,), t))
    xys = torch.cat(xs)
    tile_of = torch.cat(tiles)
    N = xys.shape[0]
    perm = torch.randperm(N, generator=g)
    xys, tile_of = xys[perm], tile_of[perm]
    dep = (torch.randint(0, 400, (N,), generator=g).float() + 1) / 16.0     # many exact ties
    rad = torch.ones(N, dtype=torch.int32)
    con = torch.tensor([[0.5, 0.0, 0.5]]).repeat(N, 1)
    rz.clear_bin_cache()
    recs, bins = rz.pack_and_bin(xys.to(DEV), dep.to(DEV), rad.to(DEV), con.to(DEV), torch.full((N,), 0.5).to(DEV),
                                 torch.rand(N, 3, generator=g).to(DEV), H, W, cull_mode=0, reuse=False)
    assert bins.num_intersects == N and bins.max_per_tile == max(sizes)
    off = bins.tile_offsets.cpu().long()
    ids = bins.ids_sorted.cpu().long()
    for t, n in enumerate(sizes):
        assert off[t + 1] - off[t] == n
        members = torch.nonzero(tile_of == t).flatten()
        want = members[torch.argsort(dep[members], stable=True)]      # members ascending -> ties by id
        assert torch.equal(ids[off[t]:off[t + 1]], want), f"tile {t} (n={n})"


def test_big_tile_fallback_sort():
    """A tile list longer than the shared-memory sort capacity goes through the global-memory
    fallback and must still be exactly sorted."""
    from tinysplat_b200 import rasterize as rz, _lib
    cap = _lib.load().ts_bin_smem_sort_cap()
    N = cap + 1500
    g = torch.Generator().manual_seed(3)
    xys = torch.rand(N, 2, generator=g) * 14 + 1          # all inside tile (0,0)
    dep = torch.rand(N, generator=g) + 0.5
    rad = torch.ones(N, dtype=torch.int32)
    con = torch.tensor([[0.5, 0.0, 0.5]]).repeat(N, 1)
    opac = torch.full((N,), 0.5)
    colors = torch.rand(N, 3, generator=g)
    xys[:, :] = xys.clamp(2, 13)
    rz.clear_bin_cache()
    recs, bins = rz.pack_and_bin(xys.to(DEV), dep.to(DEV), rad.to(DEV), con.to(DEV), opac.to(DEV),
                                 colors.to(DEV), 16, 16, cull_mode=0, reuse=False)
    assert bins.num_intersects == N and bins.max_per_tile == N
    want = torch.argsort(dep, stable=True)
    assert torch.equal(bins.ids_sorted.cpu().long(), want)


def test_full_adapter_matches_golden_config1(blend_mode):
    """BASELINE config 1 (256 Gaussians, 128x128) through both pipelines of the adapter mirror,
    against the committed fp64-oracle vectors."""
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    gold = np.load(GOLD)
    sc = {k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("in_")}
    cam = synthetic.make_camera(128, 128)
    g = torch.Generator().manual_seed(1234)
    wi = torch.rand(128, 128, 3, generator=g, dtype=torch.float64).float().to(DEV)
    wd = torch.rand(128, 128, generator=g, dtype=torch.float64).float().to(DEV)
    for pipeline in ("reference", "unfused4", "fused"):
        model = ParamModel(sc, DEV, 3)
        img, ex = GaussianRasterizer(model, None, DEV, pipeline)(cam, None, 3)
        assert torch.equal(ex["radii"].cpu(), torch.from_numpy(gold["radii"]))
        assert (img.cpu() - torch.from_numpy(gold["img"])).abs().max().item() < TOL_IMG
        assert (ex["depth"].cpu() - torch.from_numpy(gold["depth"])).abs().max().item() < 20 * TOL_IMG
        ((img * wi).sum() + 0.1 * (ex["depth"] * wd).sum()).backward()
        assert rel_err(ex["xys"].grad, torch.from_numpy(gold["v_xys"])) < TOL_GRAD
        for k in PARAMS:
            assert rel_err(getattr(model, k).grad, torch.from_numpy(gold["v_" + k])) < TOL_GRAD, (pipeline, k)


@pytest.mark.parametrize("pipeline", ["reference", "fused"])
def test_no_grad_forward_and_retain_graph(pipeline, blend_mode):
    """The viewer renders under no_grad [REF tinysplat/viewer.py:90-93]; training calls
    backward(retain_graph=True) [REF scripts/train.py:94]: a second backward must reproduce the
    first (saved buffers are neither freed nor mutated)."""
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H = 96, 64
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(500, W, H, seed=4)
    model = ParamModel(sc, DEV, 2)
    rast = GaussianRasterizer(model, None, DEV, pipeline)
    with torch.no_grad():
        img0, ex0 = rast(cam, (W, H), 2)
    assert not img0.requires_grad and ex0["xys"].grad_fn is None
    img, ex = rast(cam, (W, H), 2)
    assert torch.equal(img, img0)
    loss = img.sum() + ex["depth"].sum()
    loss.backward(retain_graph=True)
    g1 = [p.grad.clone() for p in model.parameters()]
    xg1 = ex["xys"].grad.clone()
    model.zero_grad()
    ex["xys"].grad = None
    loss.backward()
    for a, p in zip(g1, model.parameters()):
        assert rel_err(p.grad, a) < 1e-5
    assert rel_err(ex["xys"].grad, xg1) < 1e-5
    assert ex["xys"].grad.norm(dim=-1).shape == (500,)    # what update_grad_accum reads


def test_fused_node_matches_unfused_ops_on_a_larger_scene(blend_mode):
    """The single fused autograd node (activations folded into the kernels, packed gradients
    consumed in place) against the composition of the five public ops, 20k Gaussians, with a
    depth loss so the depth cotangent path (colour channel 3 -> v_depths) is live."""
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H, N = 400, 300, 20000
    cam = synthetic.make_camera(W, H, yaw_deg=5.0, shift=(0.1, 0.05, 0.0))
    sc = synthetic.make_scene(N, W, H, seed=12, sh_degree=3)
    sc["background"] = torch.tensor([0.3, 0.1, 0.7])
    g = torch.Generator().manual_seed(3)
    sc["quats"] = sc["quats"] * (0.5 + torch.rand(N, 1, generator=g))        # un-normalised on purpose
    wi = torch.rand(H, W, 3, generator=g).to(DEV)
    wd = torch.rand(H, W, generator=g).to(DEV)
    res = {}
    for pipeline in ("reference", "fused"):
        model = ParamModel(sc, DEV, 3)
        img, ex = GaussianRasterizer(model, None, DEV, pipeline)(cam, (W, H), 2)
        ((img * wi).sum() + 0.05 * (ex["depth"] * wd).sum()).backward()
        res[pipeline] = (img, ex["depth"], ex["xys"].grad, [p.grad for p in model.parameters()], ex["radii"])
    a, b = res["reference"], res["fused"]
    assert torch.equal(a[4], b[4])
    # the two pipelines round differently BEFORE the blend (exp / normalise / sigmoid in torch vs folded
    # into the kernels): xys and conics differ by an ulp, the image by a few 1e-5 (seen: 2.9e-5)
    assert (a[0] - b[0]).abs().max().item() < TOL_IMG / 4 and (a[1] - b[1]).abs().max().item() < 1e-4
    assert rel_err(b[2], a[2]) < 1e-4
    for name, ga, gb in zip(PARAMS, a[3], b[3]):
        assert rel_err(gb, ga) < 1e-4, name


def test_empty_and_all_culled_scenes(blend_mode):
    import gsplat
    W, H = 40, 24
    bg = torch.tensor([0.1, 0.6, 0.9], device=DEV)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=DEV)
    # N = 0
    img, a = gsplat.rasterize_gaussians(z(0, 2), z(0), z(0, dt=torch.int32), z(0, 3), z(0, dt=torch.int32),
                                        z(0, 3), z(0, 1), H, W, bg)
    assert torch.equal(img, bg.expand(H, W, 3)) and a.abs().max() == 0
    # everything behind the camera
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(64, W, H, seed=1)
    sc["means"][:, 2] = -2.0
    ins = proj_inputs(sc, cam, W, H, torch.float32, DEV)
    ins[0].requires_grad_(True)
    xys, dep, rad, con, nt, _ = gsplat.project_gaussians(*ins)
    assert rad.abs().sum() == 0 and nt.abs().sum() == 0
    img, a = gsplat.rasterize_gaussians(xys, dep, rad, con, nt, z(64, 3) + 0.5, z(64, 1) + 0.5, H, W, bg)
    assert torch.equal(img, bg.expand(H, W, 3))
    img.sum().backward()
    assert ins[0].grad.abs().max() == 0


def test_fused_pipeline_empty_scene_and_no_grad_outputs(blend_mode):
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H = 50, 34
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(16, W, H, seed=2)
    sc["background"] = torch.tensor([0.2, 0.4, 0.6])
    empty = {k: (v[:0] if k != "background" else v) for k, v in sc.items()}
    model = ParamModel(empty, DEV, 3)
    # poison the caching allocator: torch.empty() must not be able to hide a missing memset
    junk = [torch.full((1 << 18,), 0x7F7F7F7F, dtype=torch.int32, device=DEV) for _ in range(8)]
    del junk
    img, ex = GaussianRasterizer(model, None, DEV, "fused")(cam, (W, H), 3)
    assert torch.allclose(img, sc["background"].to(DEV).expand(H, W, 3))
    assert torch.allclose(ex["depth"], torch.full((H, W), 0.2, device=DEV))   # depth over background[0]
    img.sum().backward()
    assert all(p.grad is not None and p.grad.numel() == 0 for p in model.parameters())
    # everything behind the camera: zero grads, finite
    sc["means"][:, 2] = -1.0
    model = ParamModel(sc, DEV, 3)
    img, ex = GaussianRasterizer(model, None, DEV, "fused")(cam, (W, H), 3)
    (img.sum() + ex["depth"].sum()).backward()
    assert ex["radii"].abs().sum() == 0
    for p in model.parameters():
        assert p.grad.abs().max() == 0


@pytest.mark.parametrize("pipeline", ["reference", "fused"])
def test_huge_and_extreme_gaussians_match_oracle(pipeline, blend_mode):
    """Screen-filling Gaussians (warp-cooperative tile expansion, hundreds of tiles each), nearly
    opaque and nearly transparent ones, strongly anisotropic ones, mixed with ordinary ones."""
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H, N = 208, 144, 300
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(N, W, H, seed=31, sh_degree=2)
    sc["background"] = torch.tensor([0.9, 0.1, 0.5])
    sc["scales"][:6] += 4.5                     # enormous: cover the whole image
    sc["scales"][6:12, 0] += 3.0                # long thin needles
    sc["scales"][6:12, 1] -= 1.5
    sc["opacities"][12:20] = 9.0                # alpha clamps at 0.999
    sc["opacities"][20:30] = -6.0               # below 1/255: can never contribute
    sc["opacities"][:3] = 8.0
    names = PARAMS
    model = ParamModel(sc, DEV, 2)
    img, ex = GaussianRasterizer(model, None, DEV, pipeline)(cam, (W, H), 2)
    g = torch.Generator().manual_seed(5)
    wi, wd = torch.rand(H, W, 3, generator=g), torch.rand(H, W, generator=g)
    ((img * wi.to(DEV)).sum() + 0.1 * (ex["depth"] * wd.to(DEV)).sum()).backward()
    p = {k: v.double().clone().requires_grad_(k != "background") for k, v in sc.items()}
    rimg, rex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 2)
    ((rimg * wi.double()).sum() + 0.1 * (rex["depth"] * wd.double()).sum()).backward()
    assert (ex["radii"].cpu() - rex["radii"]).abs().max().item() <= 1
    assert ex["radii"].max().item() > 400
    assert (img.cpu().double() - rimg).abs().max().item() < TOL_IMG
    assert (ex["depth"].cpu().double() - rex["depth"]).abs().max().item() < 20 * TOL_IMG
    assert rel_err(ex["xys"].grad, rex["xys"].grad) < TOL_GRAD
    for k in names:
        assert torch.isfinite(getattr(model, k).grad).all(), k
        assert rel_err(getattr(model, k).grad, p[k].grad) < TOL_GRAD, k
    assert getattr(model, "opacities").grad[20:30].abs().max() == 0     # sub-1/255 opacity: no gradient


def test_full_size_properties_1080p():
    """BASELINE-sized run (1M Gaussians, 1080p; too big for the oracle): size-independent
    properties instead — (a) opaque background conservation: out = sum w_i c_i + T*bg, so with all
    colours == 1 and bg == 1 every pixel is exactly 1 up to fp32 rounding; (b) alpha + T == 1;
    (c) linearity of the image in colour; (d) determinism of forward; (e) finite gradients whose
    colour-gradient sum equals sum of (1 - T) for a unit cotangent."""
    import gsplat
    from tinysplat_b200 import rasterize as rz
    W, H, N = 1920, 1080, 1_000_000
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(N, W, H, seed=0)
    ins = proj_inputs(sc, cam, W, H, torch.float32, DEV)
    xys, dep, rad, con, nt, _ = gsplat.project_gaussians(*ins)
    opac = torch.sigmoid(sc["opacities"]).to(DEV)
    ones = torch.ones(N, 3, device=DEV)
    bg1 = torch.ones(3, device=DEV)
    img, alpha = gsplat.rasterize_gaussians(xys, dep, rad, con, nt, ones, opac, H, W, bg1)
    assert (img - 1).abs().max().item() < 1e-4
    g = torch.Generator().manual_seed(0)
    c1 = torch.rand(N, 3, generator=g).to(DEV).requires_grad_(True)
    c2 = torch.rand(N, 3, generator=g).to(DEV)
    bg0 = torch.zeros(3, device=DEV)
    i1, a1 = gsplat.rasterize_gaussians(xys, dep, rad, con, nt, c1, opac, H, W, bg0)
    i2, _ = gsplat.rasterize_gaussians(xys, dep, rad, con, nt, c2, opac, H, W, bg0)
    i12, _ = gsplat.rasterize_gaussians(xys, dep, rad, con, nt, 0.5 * c1 + 2.0 * c2, opac, H, W, bg0)
    assert (i12 - (0.5 * i1 + 2.0 * i2)).abs().max().item() < 1e-4
    assert torch.equal(a1, alpha)
    i1b, _ = gsplat.rasterize_gaussians(xys, dep, rad, con, nt, c1, opac, H, W, bg0)
    assert torch.equal(i1, i1b)
    i1.sum().backward()
    assert torch.isfinite(c1.grad).all()
    # d(sum img)/d c_i = sum_pixels w_i ; summed over i and channels = 3 * sum_pixels (1 - T)
    assert abs(c1.grad.sum().item() / (3 * a1.double().sum().item()) - 1) < 1e-3
    assert rz.last_stats["num_intersects"] > N


def test_blend_generations_agree_at_full_size(lib):
    """1M Gaussians at 1080p through the fused adapter with both blend-backward kernels: the
    gradients must agree up to the order of the float sums — on real hardware and at full size
    this is the check that the exact per-row culling of the grouped kernel never drops a
    contribution the first-generation kernel (bounding-box culling) keeps."""
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H, N = 1920, 1080, 1_000_000
    cam = synthetic.make_camera(W, H, yaw_deg=2.0)
    sc = synthetic.make_scene(N, W, H, seed=0)
    sc["background"] = torch.tensor([0.1, 0.2, 0.3])
    g = torch.Generator().manual_seed(11)
    wi = torch.rand(H, W, 3, generator=g).to(DEV)
    wd = torch.rand(H, W, generator=g).to(DEV)
    res = {}
    try:
        for mode in (0, 1):
            assert lib.ts_set_blend_mode(mode) == 0
            model = ParamModel(sc, DEV, 3)
            img, ex = GaussianRasterizer(model, None, DEV, "fused")(cam, (W, H), 3)
            ((img * wi).sum() + 0.01 * (ex["depth"] * wd).sum()).backward()
            res[mode] = (img, ex["depth"], ex["xys"].grad, [p.grad for p in model.parameters()])
    finally:
        lib.ts_set_blend_mode(-1)
    for mode in (1,):
        assert torch.equal(res[0][0], res[mode][0]), f"image differs in mode {mode}"
        assert torch.equal(res[0][1], res[mode][1]), f"depth differs in mode {mode}"
        assert rel_err(res[mode][2], res[0][2]) < 1e-4
        for name, ga, gb in zip(PARAMS, res[0][3], res[mode][3]):
            assert torch.isfinite(gb).all(), name
            assert rel_err(gb, ga) < 1e-4, (mode, name)


class _LoopbackExchange:
    """Single-process stand-in for parallel.PackedGradExchange (world 1): the 'all-to-all' and the
    'all-gather' are identities, and every packed send buffer + camera row is recorded so that the
    multi-view shard kernels can be driven by hand afterwards."""
    world, rank, average = 1, 0, False

    def __init__(self):
        self.sent, self.cams = [], []

    def shard_rows(self, n):
        return max(128, (n + 127) // 128 * 128)

    def out_scale(self):
        return 1.0

    def buffer(self, key, shape, dtype, device, zero=False, tag=None):
        return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)

    def gather_cameras(self, row):
        self.cams.append(row.clone())
        return row.view(1, -1)

    def all_to_all_rows(self, send):
        self.sent.append(send.clone())
        return send.view(1, send.shape[0], send.shape[1])

    def all_gather_shards(self, shards):
        return shards


def test_packed_exchange_shard_kernels_match_the_plain_backward(lib):
    """SURVEY 8e, packed-row gradient exchange: (1) with a loopback exchange the fused node's
    packed backward (ts_dp_prepare -> ts_project_bwd_views / ts_sh_bwd_views over one view) must
    reproduce the plain backward; (2) the shard kernels driven with TWO views' recorded rows must
    equal the sum of the two plain backwards (what the all-reduce strategy delivers)."""
    from tinysplat_b200 import _lib
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H, N = 320, 208, 3001                      # N not a multiple of the 128-row shard unit
    sc = synthetic.make_scene(N, W, H, seed=21, sh_degree=3)
    sc["background"] = torch.tensor([0.2, 0.4, 0.1])
    sc["means"][:40, 2] = -1.0                    # culled Gaussians: their packed rows must be inert
    sc["opacities"][40:60] = 9.0
    cams = [synthetic.make_camera(W, H, yaw_deg=-4.0, shift=(0.1, 0.0, 0.0)),
            synthetic.make_camera(W, H, yaw_deg=6.0, shift=(-0.2, 0.05, 0.1))]
    g = torch.Generator().manual_seed(3)
    wi = torch.rand(H, W, 3, generator=g).to(DEV)
    wd = torch.rand(H, W, generator=g).to(DEV)

    def run(cam, exchange):
        model = ParamModel(sc, DEV, 3)
        rast = GaussianRasterizer(model, None, DEV, "fused")
        rast.grad_exchange = exchange
        img, ex = rast(cam, (W, H), 3)
        ((img * wi).sum() + 0.05 * (ex["depth"] * wd).sum()).backward()
        return model, ex["xys"].grad

    plain = [run(c, None) for c in cams]
    loop = _LoopbackExchange()
    packed = [run(c, loop) for c in cams]
    for (mp_, xg_p), (mq, xg_q) in zip(plain, packed):
        assert rel_err(xg_q, xg_p) < 1e-5
        for name in PARAMS:
            a, b = getattr(mq, name).grad, getattr(mp_, name).grad
            assert a.shape == b.shape and torch.isfinite(a).all(), name
            assert rel_err(a, b) < 1e-5, name

    # (2) two views at once through the C ABI
    Ns = loop.shard_rows(N)
    rows = torch.stack(loop.sent).contiguous()                   # [2, Ns, 12]
    cam_rows = torch.stack(loop.cams).contiguous()               # [2, 32]
    m = plain[0][0]
    K = m.colors_rest.shape[1] + 1
    f32 = dict(device=DEV, dtype=torch.float32)
    v_means, v_scales, v_quats, v_logit = (torch.empty(N, 3, **f32), torch.empty(N, 3, **f32),
                                           torch.empty(N, 4, **f32), torch.empty(N, **f32))
    v_dc, v_rest = torch.empty(N, 1, 3, **f32), torch.empty(N, K - 1, 3, **f32)
    st = _lib.stream_ptr(torch.device(DEV))
    flags = _lib.PROJ_LOG_SCALES | _lib.PROJ_RAW_QUATS | _lib.PROJ_DEPTH_CH3
    _lib.call("ts_project_bwd_views", 2, N, _lib.ptr(m.means.detach()), _lib.ptr(m.scales.detach()), 1.0,
              _lib.ptr(m.quats.detach()), _lib.ptr(cam_rows), H, W, flags, _lib.ptr(rows), Ns * 12,
              _lib.ptr(m.opacities.detach().reshape(-1)), 0.5, _lib.ptr(v_means), _lib.ptr(v_scales),
              _lib.ptr(v_quats), _lib.ptr(v_logit), st)
    _lib.call("ts_sh_bwd_views", 2, N, 3, K, _lib.ptr(m.means.detach()), _lib.ptr(cam_rows), _lib.ptr(rows),
              Ns * 12, 0.5, _lib.ptr(v_dc), _lib.ptr(v_rest), st)
    got = dict(means=v_means, scales=v_scales, quats=v_quats, opacities=v_logit.reshape(N, 1),
               colors_dc=v_dc.reshape(m.colors_dc.shape), colors_rest=v_rest)
    for name in PARAMS:
        want = 0.5 * (getattr(plain[0][0], name).grad + getattr(plain[1][0], name).grad)
        assert rel_err(got[name], want) < 1e-5, name
    assert got["means"][:40].abs().max() == 0


def test_peer_exchange_single_rank_matches_the_plain_backward(lib):
    """SURVEY 8e, peer-memory exchange (csrc/peer.cu, parallel.PeerGradExchange) on ONE GPU: a
    single-rank process group, so every 'peer' is this device, but the whole mechanism runs on the
    hardware — cudaMalloc'ed exchange buffer aliased by torch views, ts_dp_push, the release/acquire
    flag barrier, SH-backward over colour rows, shard projection-backward storing through pointer
    tables — and must reproduce the plain backward.  Multi-rank: tools/dp_check.py (2 and 8 GPUs),
    tests/test_pipeline_emu.py (2-4 simulated ranks on the emulator)."""
    import torch.distributed as dist
    from tinysplat_b200.parallel import PeerGradExchange
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    if dist.is_initialized():
        pytest.skip("a process group already exists")
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1,
                            device_id=torch.device(DEV))
    try:
        W, H = 320, 208
        g = torch.Generator().manual_seed(3)
        wi = torch.rand(H, W, 3, generator=g).to(DEV)
        wd = torch.rand(H, W, generator=g).to(DEV)
        ex_peer = PeerGradExchange(average=True)
        for N in (3001, 2000, 5000):              # shrink within the capacity, then grow past it (re-allocation)
            sc = synthetic.make_scene(N, W, H, seed=21, sh_degree=3)
            sc["background"] = torch.tensor([0.2, 0.4, 0.1])
            sc["means"][:40, 2] = -1.0
            cam = synthetic.make_camera(W, H, yaw_deg=-4.0, shift=(0.1, 0.0, 0.0))
            got = {}
            for name, exch in (("plain", None), ("peer", ex_peer)):
                model = ParamModel(sc, DEV, 3)
                rast = GaussianRasterizer(model, None, DEV, "fused")
                rast.grad_exchange = exch
                for _ in range(2):                # twice: the second step reuses the buffers and a new epoch
                    model.zero_grad()
                    img, ex = rast(cam, (W, H), 3)
                    ((img * wi).sum() + 0.05 * (ex["depth"] * wd).sum()).backward()
                got[name] = ({k: getattr(model, k).grad.clone() for k in PARAMS}, ex["xys"].grad.clone())
            ex_peer.check()
            assert rel_err(got["peer"][1], got["plain"][1]) < 1e-5
            for k in PARAMS:
                a, b = got["peer"][0][k], got["plain"][0][k]
                assert a.shape == b.shape and torch.isfinite(a).all(), k
                assert rel_err(a, b) < 1e-5, (N, k)
            assert got["peer"][0]["means"][:40].abs().max() == 0
        ex_peer.close()
    finally:
        dist.destroy_process_group()


# ---- fused Adam ---------------------------------------------------------------------
def _param_set(N, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = {"means": (N, 3), "colors_dc": (N, 3), "colors_rest": (N, 15, 3), "scales": (N, 3),
              "quats": (N, 4), "opacities": (N, 1)}
    lrs = {"means": 0.00016, "colors_dc": 0.0025, "colors_rest": 0.000125, "scales": 0.005,
           "quats": 0.001, "opacities": 0.05}          # [REF scripts/train.py:183-188]
    return {k: torch.randn(*s, generator=g) for k, s in shapes.items()}, lrs


def _groups(ps, lrs):
    return [{"params": [p], "lr": lrs[k], "name": k} for k, p in ps.items()]


def _surgery(optim, ps, mask, n_new, seed):
    """The reference's densify/prune optimizer surgery [REF model_gaussian.py:199-242], restated."""
    g = torch.Generator().manual_seed(seed)
    for group in optim.param_groups:
        name = group["name"]
        old = group["params"][0]
        new_rows = torch.randn(n_new, *old.shape[1:], generator=g).to(old.device)
        state = optim.state[old]
        state["exp_avg"] = torch.cat((state["exp_avg"][~mask], torch.zeros_like(new_rows)))
        state["exp_avg_sq"] = torch.cat((state["exp_avg_sq"][~mask], torch.zeros_like(new_rows)))
        del optim.state[old]
        new = torch.nn.Parameter(torch.cat((old.detach()[~mask], new_rows)))
        group["params"][0] = new
        optim.state[new] = state
        ps[name] = new


def test_fused_adam_matches_torch_adam_through_optimizer_surgery():
    from tinysplat_b200.optim import FusedAdam
    N = 5003                                   # odd sizes: exercises the scalar tails
    init, lrs = _param_set(N, 0)
    pa = {k: torch.nn.Parameter(v.clone().to(DEV)) for k, v in init.items()}
    pb = {k: torch.nn.Parameter(v.clone().to(DEV)) for k, v in init.items()}
    ref = torch.optim.Adam(_groups(pa, lrs), foreach=False, fused=False)
    ours = FusedAdam(_groups(pb, lrs))
    g = torch.Generator().manual_seed(1)

    def step_both(n_steps):
        for _ in range(n_steps):
            for k in pa:
                gr = torch.randn(*pa[k].shape, generator=g).to(DEV) * (10.0 ** float(torch.randint(-3, 2, (1,), generator=g)))
                pa[k].grad = gr.clone()
                pb[k].grad = gr.clone()
            ref.step()
            ours.step()
            for k in pa:
                # fp32 vs fp32 with possibly different FMA contraction: a few ulp of the tensor's scale
                sa, sb = ref.state[pa[k]], ours.state[pb[k]]
                assert float(sa["step"]) == float(sb["step"])
                errs = (rel_err(pb[k], pa[k]), rel_err(sb["exp_avg"], sa["exp_avg"]),
                        rel_err(sb["exp_avg_sq"], sa["exp_avg_sq"]))
                assert max(errs) < 2e-6, (k, errs)

    step_both(4)
    mask = torch.rand(N, generator=g).to(DEV) < 0.2
    _surgery(ref, pa, mask, 700, seed=5)
    _surgery(ours, pb, mask, 700, seed=5)
    step_both(3)
    assert pa["means"].shape[0] == N - int(mask.sum()) + 700


def test_fused_adam_rejects_unsupported_options():
    from tinysplat_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(8, 3, device=DEV))
    opt = FusedAdam([{"params": [p], "lr": 0.1, "name": "x"}])
    opt.param_groups[0]["weight_decay"] = 0.1
    p.grad = torch.ones_like(p)
    with pytest.raises(NotImplementedError):
        opt.step()


def test_training_loop_like_train_py_reduces_the_loss():
    """The loop of scripts/train.py [REF train.py:45-97] on a synthetic target: render, loss =
    0.8 L1 + 0.2 (1 - SSIM), backward, Adam — with the fused adapter, fused SSIM and FusedAdam.
    Gradients that were wrong in sign or scale would not bring the loss down."""
    from tinysplat_b200.optim import FusedAdam
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    from tinysplat_b200.ssim import SSIM
    W, H, N = 112, 96, 1500
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(N, W, H, seed=8, sh_degree=3)
    sc["background"] = torch.tensor([0.1, 0.1, 0.1])
    with torch.no_grad():
        target, _ = GaussianRasterizer(ParamModel(sc, DEV, 3, requires_grad=False), None, DEV)(cam, None, 3)
    g = torch.Generator().manual_seed(0)
    start = dict(sc)
    start["colors_dc"] = sc["colors_dc"] + 0.5 * torch.randn(N, 3, generator=g)
    start["opacities"] = sc["opacities"] + 0.5 * torch.randn(N, 1, generator=g)
    start["means"] = sc["means"] + 0.01 * torch.randn(N, 3, generator=g)
    model = ParamModel(start, DEV, 3)
    lrs = {"means": 0.0005, "colors_dc": 0.03, "colors_rest": 0.002, "scales": 0.005, "quats": 0.001, "opacities": 0.05}
    names = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]
    params = {k: torch.nn.Parameter(getattr(model, k).detach()) for k in names}
    for k, p in params.items():
        setattr(model, k, p)
    opt = FusedAdam([{"params": [params[k]], "lr": lrs[k], "name": k} for k in names])
    rast = GaussianRasterizer(model, None, DEV)
    ssim = SSIM(data_range=1.0, size_average=True, channel=3)
    losses = []
    for step in range(60):
        img, extras = rast(cam, None, 3)
        l1 = (img - target).abs().mean()
        dssim = 1 - ssim(img.permute(2, 0, 1).unsqueeze(0), target.permute(2, 0, 1).unsqueeze(0))
        loss = 0.8 * l1 + 0.2 * dssim
        loss.backward(retain_graph=True)            # as train.py does
        opt.step()
        assert extras["xys"].grad is not None and extras["xys"].grad.norm(dim=-1).shape == (N,)
        opt.zero_grad(set_to_none=True)
        losses.append(float(loss))
    assert all(torch.isfinite(p).all() for p in params.values())
    assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])


@pytest.mark.parametrize("P1,P2,K", [(500, 5000, 16), (33, 1500, 4), (700, 20, 16), (1, 1, 1), (10, 5, 16)])
def test_knn_points_matches_brute_force(P1, P2, K):
    """exact KNN as the density regularizer calls it [REF model_gaussian.py:260]."""
    from tinysplat_b200.knn import knn_points
    g = torch.Generator().manual_seed(P1 + P2)
    p1 = torch.randn(P1, 3, generator=g)
    p2 = torch.randn(P2, 3, generator=g)
    out = knn_points(p1[None].to(DEV), p2[None].to(DEV), K=K)
    assert out.idx.shape == (1, P1, K) and out.idx.dtype == torch.int64 and out.dists.shape == (1, P1, K)
    d = ((p1.double()[:, None, :] - p2.double()[None, :, :]) ** 2).sum(-1)
    kk = min(K, P2)
    want_d, want_i = torch.topk(d, kk, dim=1, largest=False, sorted=True)
    got_d, got_i = out.dists[0].cpu().double()[:, :kk], out.idx[0].cpu()[:, :kk]
    assert torch.allclose(got_d, want_d, rtol=1e-5, atol=1e-6)
    assert (got_i == want_i).float().mean().item() > 0.999        # fp32 near-ties may swap neighbours
    # what the reference does with it: index the means
    assert p2.to(DEV)[out.idx[0]].shape == (P1, K, 3)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pipeline", ["reference", "fused"])
def test_guessed_buffer_sizes_never_change_the_result(pipeline):
    """emit / sort / blend are queued before the host knows the intersection count, with buffers sized
    from earlier calls (tinysplat_b200/binning.py).  A sparse scene first, then a much denser one at the
    same image size (the guess is too small -> the pass is repeated), then the dense one again (the
    guess fits): every render must equal the one made with exact sizes."""
    from tinysplat_b200 import binning, rasterize as rz
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H = 208, 112
    cam = synthetic.make_camera(W, H)
    sparse = synthetic.make_scene(300, W, H, seed=1, mean_radius_px=3.0)
    dense = synthetic.make_scene(6000, W, H, seed=2, mean_radius_px=14.0)

    def render(sc):
        model = ParamModel(sc, DEV, 3)
        rz.clear_bin_cache()
        img, ex = GaussianRasterizer(model, None, DEV, pipeline)(cam, (W, H), 3)
        (img.sum() + 0.1 * ex["depth"].sum()).backward()
        return img.detach().clone(), ex["depth"].detach().clone(), [p.grad.clone() for p in model.parameters()], \
            rz.last_stats["num_intersects"]

    binning.reset_state()
    exact = render(dense)                      # first call of this image size: sized exactly
    binning.reset_state()
    before = dict(binning.stats)
    render(sparse)                             # seeds a small guess
    redone = render(dense)                     # guess too small: repeated with exact sizes
    assert binning.stats["redone"] > before["redone"]
    n_redone = binning.stats["redone"]
    again = render(dense)                      # guess fits now
    assert binning.stats["redone"] == n_redone and binning.stats["speculative"] > before["speculative"]
    for got in (redone, again):
        assert got[3] == exact[3]
        assert torch.equal(got[0], exact[0]) and torch.equal(got[1], exact[1])
        for a, b in zip(got[2], exact[2]):
            assert rel_err(a, b) < 1e-5        # float atomics reorder the sums


def test_depth_pass_reuses_the_tile_lists_of_the_rgb_pass():
    """The reference rasterises RGB and then depth over the same geometry, building a FRESH
    torch.sigmoid(model.opacities) for each call [REF rasterize.py:44-51,86]: the second call must find
    the first call's tile lists although its opacity tensor is a new one."""
    import gsplat
    from gsplat.sh import spherical_harmonics
    from tinysplat_b200 import rasterize as rz
    from tinysplat_b200.rasterizer import ParamModel
    W, H = 160, 96
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(2000, W, H, seed=9)
    model = ParamModel(sc, DEV, 3)
    xys, depths, radii, conics, num_tiles, _ = gsplat.project_gaussians(*proj_inputs(
        {k: getattr(model, k) for k in ("means", "scales", "quats")}, cam, W, H, device=DEV))
    dirs = torch.nn.functional.normalize(model.means - cam.view_matrix[:3, 3].to(DEV), dim=-1)
    rgbs = torch.clamp(spherical_harmonics(3, dirs, torch.cat([model.colors_dc[:, None], model.colors_rest], 1)) + 0.5, min=0)
    rz.clear_bin_cache()
    bg = torch.zeros(3, device=DEV)
    img, _ = gsplat.rasterize_gaussians(xys, depths, radii, conics, num_tiles, rgbs, torch.sigmoid(model.opacities), H, W, bg)
    assert rz.last_stats["bins_reused"] is False
    dimg, _ = gsplat.rasterize_gaussians(xys, depths, radii, conics, num_tiles, depths[:, None].repeat(1, 3),
                                         torch.sigmoid(model.opacities), H, W, bg)
    assert rz.last_stats["bins_reused"] is True
    # a changed opacity must NOT reuse them (the footprint culling depends on it)
    with torch.no_grad():
        model.opacities.add_(0.5)
    img2, _ = gsplat.rasterize_gaussians(xys, depths, radii, conics, num_tiles, rgbs, torch.sigmoid(model.opacities), H, W, bg)
    assert rz.last_stats["bins_reused"] is False
    assert (img2 - img).abs().max().item() > 0


def test_device_approximations_the_blend_kernels_rely_on(lib):
    """Blend-backward multiplies every row's transmittance by rcp(1 - alpha) with alpha = 0 where the
    Gaussian does not reach the pixel: MUFU.RCP(1) must be exactly 1 (and stay within an ulp or two
    elsewhere); MUFU.EX2 within 2 ulp + the documented 2^-22 relative error."""
    from tinysplat_b200 import _lib
    x = torch.cat([torch.tensor([1.0, 2.0, 0.5, 0.25, 4.0]), torch.linspace(0.001, 1.0, 4000)]).to(DEV)
    rcp, ex2 = torch.empty_like(x), torch.empty_like(x)
    _lib.call("ts_debug_approx", x.numel(), _lib.ptr(x), _lib.ptr(rcp), _lib.ptr(ex2), _lib.stream_ptr(x.device))
    torch.cuda.synchronize()
    assert rcp[0].item() == 1.0 and rcp[1].item() == 0.5 and rcp[2].item() == 2.0
    assert ((rcp - 1.0 / x.double()).abs() / (1.0 / x.double())).max().item() < 2.5e-7
    assert ((ex2 - torch.exp2(x.double())).abs() / torch.exp2(x.double())).max().item() < 5e-7


@pytest.mark.parametrize("ch", [1, 3, 4])
def test_tile_launch_order_does_not_change_the_result(lib, ch):
    """The blend kernels take their tiles longest-list-first (ts_bin_tile_order); image, alpha and the
    gradients must equal the raster-order launch (forward bit for bit, backward up to the order of the
    float atomics)."""
    import gsplat
    from tinysplat_b200 import binning, rasterize as rz
    N, W, H = 30000, 500, 300
    xys, dep, rad, con, nt, colors, opac, bg = _raster_case(N, W, H, seed=5, CH=ch)
    res = []
    for use in (True, False):
        binning.USE_TILE_ORDER = use
        try:
            rz.clear_bin_cache()
            leaves = [t.to(DEV).requires_grad_(True) for t in (xys, con, colors, opac)]
            img, alpha = gsplat.rasterize_gaussians(leaves[0], dep.to(DEV), rad.to(DEV), leaves[1], nt.to(DEV),
                                                    leaves[2], leaves[3], H, W, bg.to(DEV))
            (img.square().sum() + alpha.sum()).backward()
            res.append((img.detach(), alpha.detach(), [l.grad for l in leaves]))
        finally:
            binning.USE_TILE_ORDER = True
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for a, b in zip(res[0][2], res[1][2]):
        assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("shape,u8", [((1080, 1920, 3), True), ((1080, 1920, 3), False), ((37, 21, 3), True), ((5,), False)])
def test_fused_l1_loss_matches_torch(shape, u8):
    """tinysplat_b200.loss.l1_loss (one pass: value + gradient; uint8 ground truth = value / 255) against
    the reference's expression `(rendered - gt).abs().mean()` [REF scripts/train.py:58-59]."""
    from tinysplat_b200.loss import l1_loss
    g = torch.Generator().manual_seed(5)
    img = torch.rand(*shape, generator=g).to(DEV).requires_grad_(True)
    raw = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8).to(DEV) if u8 else torch.rand(*shape, generator=g).to(DEV)
    gt = raw / 255 if u8 else raw                       # what the reference uploads
    want = (img - gt).abs().mean()
    (want * 3.0).backward()
    want_grad, img.grad = img.grad.clone(), None
    for _ in range(2):                                  # twice: the kernel's block counter must reset itself
        got = l1_loss(img, raw)
        (got * 3.0).backward()
        assert abs(got.item() - want.item()) < 1e-6
        assert torch.equal(img.grad, want_grad)
        img.grad = None
    with torch.no_grad():
        assert abs(l1_loss(img, raw).item() - want.item()) < 1e-6


def test_fused_backward_tail_equals_the_two_kernel_tail():
    """ts_project_sh_bwd (projection-backward + SH-backward in one kernel, the single-GPU default) against
    ts_project_bwd and ts_sh_bwd on two streams: the same arithmetic, so the same bits."""
    from tinysplat_b200 import fused
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    W, H, N = 320, 200, 70001                      # a ragged last block
    cam = synthetic.make_camera(W, H, yaw_deg=2.0)
    sc = synthetic.make_scene(N, W, H, seed=21, sh_degree=3)
    g = torch.Generator().manual_seed(4)
    wi, wd = torch.rand(H, W, 3, generator=g).to(DEV), torch.rand(H, W, generator=g).to(DEV)
    res = []
    for deg in (3, 1):
        for tail in (True, False):
            fused.FUSED_TAIL = tail
            try:
                model = ParamModel(sc, DEV, 3)
                img, ex = GaussianRasterizer(model, None, DEV, "fused")(cam, (W, H), deg)
                ((img * wi).sum() + 0.1 * (ex["depth"] * wd).sum()).backward()
                res.append([p.grad.clone() for p in model.parameters()] + [ex["xys"].grad.clone()])
            finally:
                fused.FUSED_TAIL = True
        # blend-backward's float atomics reorder sums between the two runs: compare to their noise level
        for a, b in zip(res[-2], res[-1]):
            assert rel_err(a, b) < 1e-5
