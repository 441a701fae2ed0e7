"""Kernel-logic parity WITHOUT a GPU for K3 (tinysplat_b200/csrc/binning.cu: tile count -> look-back
scan -> emit -> per-tile sort), compiled as host code on the fiber SIMT emulator in tests/emu and
compared with the oracle's bin_and_sort.  Integer / index work: bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from oracle import gsplat_oracle as go
from tinysplat_b200 import synthetic

import emu_lib


@pytest.fixture(scope="module")
def emu():
    return emu_lib.load()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bin(emu, xys, depths, radii, conics, opac, W, H, cull):
    n = xys.shape[0]
    tx, ty = (W + 15) // 16, (H + 15) // 16
    T = tx * ty
    stride = emu.emu_bin_counter_stride()
    xs = np.ascontiguousarray(xys.numpy().astype(np.float32))
    dp = np.ascontiguousarray(depths.numpy().astype(np.float32))
    rd = np.ascontiguousarray(radii.numpy().astype(np.int32))
    cn = np.ascontiguousarray(conics.numpy().astype(np.float32))
    op = np.ascontiguousarray(opac.reshape(-1).numpy().astype(np.float32))
    col = np.zeros((n, 3), dtype=np.float32)
    recs = np.zeros((n, 12), dtype=np.float32)
    counts = np.full(T * stride, -1, dtype=np.int32)
    assert emu.emu_bin_count(n, 3, _ptr(xs), _ptr(rd), _ptr(cn), _ptr(op), _ptr(col), tx, ty, cull, 0,
                             _ptr(recs), _ptr(counts)) == 0
    per_tile = counts[::stride].copy()
    offsets = np.zeros(T + 1, dtype=np.int32)
    stats = np.zeros(emu.emu_bin_scan_work_ints(), dtype=np.int32)
    assert emu.emu_bin_scan(T, _ptr(counts), _ptr(offsets), _ptr(stats), emu.emu_bin_smem_sort_cap()) == 0
    M, max_count, n_big = int(stats[0]), int(stats[1]), int(stats[2])
    assert np.array_equal(offsets[1:], np.cumsum(per_tile)) and offsets[0] == 0 and M == per_tile.sum()
    assert max_count == (per_tile.max() if T else 0)
    assert np.array_equal(counts[::stride], offsets[:-1])      # counters rewritten as emit cursors
    keys = np.zeros(max(M, 1), dtype=np.uint64)
    ids = np.full(max(M, 1), -1, dtype=np.int32)
    assert emu.emu_bin_emit(n, _ptr(dp), _ptr(rd), _ptr(recs), tx, ty, cull, _ptr(counts), _ptr(keys)) == 0
    assert np.array_equal(counts[::stride], offsets[1:])       # every cursor ends at the next offset
    scratch = cnt = None
    if n_big:
        P = 1 << (max_count - 1).bit_length()
        scratch, cnt = np.zeros(n_big * P, dtype=np.uint64), np.zeros(1, dtype=np.int32)
    assert emu.emu_bin_sort(T, _ptr(offsets), _ptr(keys), _ptr(ids), max_count, n_big, _ptr(scratch), _ptr(cnt)) == 0
    return offsets, ids[:M], recs, (tx, ty)


def _project(n, W, H, seed, radius):
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(n, W, H, seed=seed, mean_radius_px=radius)
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    with torch.no_grad():
        xys, depths, radii, conics, ntiles, _ = oracle.project_gaussians(
            sc["means"], sc["scales"].exp(), 1.0, torch.nn.functional.normalize(sc["quats"], dim=-1),
            cam.view_matrix[:3], cam.proj_matrix @ cam.view_matrix, cam.f_x, cam.f_y, W / 2, H / 2, H, W, tb)
    return xys, depths, radii, conics, torch.sigmoid(sc["opacities"]), tb


@pytest.mark.parametrize("n,W,H,radius", [
    (600, 80, 56, 5.0),       # ordinary: warp-per-tile register sort, ragged tile grid
    (900, 16, 16, 3.0),       # one tile, 513..2048 entries: CTA shared-memory sort, first size class
    (2600, 16, 16, 3.0),      # one tile, > 2048 entries: second size class
    (40, 128, 96, 60.0),      # screen-filling Gaussians: warp-cooperative tile expansion
])
def test_binning_kernels_reproduce_the_oracle_lists(emu, n, W, H, radius):
    xys, depths, radii, conics, opac, tb = _project(n, W, H, 11, radius)
    offsets, ids, recs, _ = _bin(emu, xys, depths, radii, conics, opac, W, H, cull=0)
    tile, gid = go.bin_and_sort(xys, depths, radii, tb)
    T = tb[0] * tb[1]
    want_off = np.zeros(T + 1, dtype=np.int64)
    want_off[1:] = np.cumsum(np.bincount(tile.numpy(), minlength=T))
    assert np.array_equal(offsets, want_off.astype(np.int32))
    assert np.array_equal(ids, gid.numpy().astype(np.int32))       # same order: depth, then id


def test_footprint_culling_keeps_a_sorted_superset_of_the_lit_pairs(emu):
    n, W, H = 500, 96, 64
    xys, depths, radii, conics, opac, tb = _project(n, W, H, 4, 6.0)
    opac[::5] = 0.001                                  # below 1/255: never emitted
    off0, ids0, _, _ = _bin(emu, xys, depths, radii, conics, opac, W, H, cull=0)
    off1, ids1, _, _ = _bin(emu, xys, depths, radii, conics, opac, W, H, cull=1)
    assert off1[-1] < off0[-1]
    T = tb[0] * tb[1]
    d = depths.numpy()
    for t in range(T):
        full = ids0[off0[t]:off0[t + 1]]
        kept = ids1[off1[t]:off1[t + 1]]
        assert set(kept.tolist()) <= set(full.tolist())
        assert np.array_equal(kept, np.array([g for g in full if g in set(kept.tolist())], dtype=np.int32))
        # every dropped pair is really dark in this tile: alpha < 1/255 at all of its pixel centres
        ty_, tx_ = divmod(t, tb[0])
        px = tx_ * 16 + np.arange(16) + 0.5
        py = ty_ * 16 + np.arange(16) + 0.5
        for g in set(full.tolist()) - set(kept.tolist()):
            dx = xys[g, 0].item() - px[None, :]
            dy = xys[g, 1].item() - py[:, None]
            a, b, c = conics[g].tolist()
            sig = 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy
            alpha = opac[g].item() * np.exp(-sig)
            assert (alpha[sig >= 0] < 1.0 / 255.0).all(), (t, g)
    assert not np.isin(ids1, np.arange(0, n, 5)).any()
    assert d is not None


def _sort_lists(emu, sizes, seed, max_count=None, cap=0):
    """ts_bin_sort alone over hand-made tile lists: keys = depth bits << 32 | id with many depth ties."""
    rng = np.random.default_rng(seed)
    offsets = np.zeros(len(sizes) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(sizes)
    M = int(offsets[-1])
    depth = (rng.integers(1, 300, size=M).astype(np.float32) / 8.0).view(np.uint32).astype(np.uint64)
    ids_in = rng.permutation(M).astype(np.uint64)
    keys = (depth << np.uint64(32)) | ids_in
    ids = np.full(M, -7, dtype=np.int32)
    mc = max(sizes) if max_count is None else max_count
    emu.emu_set_key_capacity(cap)
    try:
        assert emu.emu_bin_sort(len(sizes), _ptr(offsets), _ptr(keys.copy()), _ptr(ids), mc, 0, None, None) == 0
    finally:
        emu.emu_set_key_capacity(0)
    return offsets, keys, ids


@pytest.mark.parametrize("sizes", [
    [513, 1024, 1025, 0, 2047, 2048, 5, 700],          # register chunk sort + merge levels, 256 threads x 4 / 8 keys
    [2049, 4096, 4097, 3, 8192, 600],                  # 512 threads x 8 / 16 keys
    [8193, 16384, 100, 12345],                         # 1024 threads x 16 keys (the shared-memory capacity)
])
def test_merge_sort_size_classes(emu, sizes):
    offsets, keys, ids = _sort_lists(emu, sizes, seed=len(sizes))
    for t, n in enumerate(sizes):
        k = keys[offsets[t]:offsets[t + 1]]
        want = (np.sort(k) & np.uint64(0xffffffff)).astype(np.int32)     # unique keys: depth, then id
        assert np.array_equal(ids[offsets[t]:offsets[t + 1]], want), f"tile {t} (n={n})"


def test_sort_with_guessed_bounds_leaves_valid_ids(emu):
    """The host launches sort before it knows M and the longest list (tinysplat_b200/binning.py).  A
    list longer than the guessed bound stays unsorted but holds its own ids (the blend kernel queued
    behind it must not gather through garbage); a list that crosses the capacity is not touched."""
    sizes = [40, 900, 300, 3000, 64]
    offsets, keys, ids = _sort_lists(emu, sizes, seed=3, max_count=1000, cap=int(np.sum(sizes)) - 30)
    low = lambda k: (k & np.uint64(0xffffffff)).astype(np.int32)
    for t in (0, 1, 2):                                                   # within the bound: sorted
        k = keys[offsets[t]:offsets[t + 1]]
        assert np.array_equal(ids[offsets[t]:offsets[t + 1]], low(np.sort(k)))
    k = keys[offsets[3]:offsets[4]]                                       # 3000 > bound: unsorted, valid
    assert np.array_equal(ids[offsets[3]:offsets[4]], low(k))
    assert (ids[offsets[4]:] == -7).all()                                 # crosses the capacity: untouched


def test_emit_respects_capacity_and_can_be_repeated(emu):
    """Emit with a too-small key buffer writes nothing past it; after ts_bin_reset_cursors the pass is
    repeated with the exact size and gives the oracle's lists."""
    n, W, H = 700, 64, 48
    xys, depths, radii, conics, opac, tb = _project(n, W, H, 5, 6.0)
    tx, ty = tb[0], tb[1]
    T = tx * ty
    stride = emu.emu_bin_counter_stride()
    xs, dp = xys.numpy().astype(np.float32), depths.numpy().astype(np.float32)
    rd, cn = radii.numpy().astype(np.int32), conics.numpy().astype(np.float32)
    op = opac.reshape(-1).numpy().astype(np.float32)
    col, recs = np.zeros((n, 3), np.float32), np.zeros((n, 12), np.float32)
    counts = np.zeros(T * stride, np.int32)
    assert emu.emu_bin_count(n, 3, _ptr(xs), _ptr(rd), _ptr(cn), _ptr(op), _ptr(col), tx, ty, 0, 0, _ptr(recs), _ptr(counts)) == 0
    offsets, stats = np.zeros(T + 1, np.int32), np.zeros(emu.emu_bin_scan_work_ints(), np.int32)
    assert emu.emu_bin_scan(T, _ptr(counts), _ptr(offsets), _ptr(stats), emu.emu_bin_smem_sort_cap()) == 0
    M, max_count = int(stats[0]), int(stats[1])
    cap = M // 2
    sentinel = np.uint64(0xdeadbeefdeadbeef)
    keys = np.full(M, sentinel, dtype=np.uint64)
    emu.emu_set_key_capacity(cap)
    try:
        assert emu.emu_bin_emit(n, _ptr(dp), _ptr(rd), _ptr(recs), tx, ty, 0, _ptr(counts), _ptr(keys)) == 0
    finally:
        emu.emu_set_key_capacity(0)
    assert (keys[cap:] == sentinel).all() and (keys[:cap] != sentinel).all()
    assert emu.emu_bin_reset_cursors(T, _ptr(offsets), _ptr(counts)) == 0
    assert np.array_equal(counts[::stride], offsets[:-1])
    ids = np.full(M, -1, np.int32)
    assert emu.emu_bin_emit(n, _ptr(dp), _ptr(rd), _ptr(recs), tx, ty, 0, _ptr(counts), _ptr(keys)) == 0
    assert emu.emu_bin_sort(T, _ptr(offsets), _ptr(keys), _ptr(ids), max_count, 0, None, None) == 0
    tile, gid = go.bin_and_sort(xys, depths, radii, tb)
    assert np.array_equal(ids, gid.numpy().astype(np.int32))
