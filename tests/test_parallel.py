"""Host logic of the data-parallel path on CPU: world_size 2, gloo.  A pure-torch stand-in
renderer replaces the CUDA rasterizer (the collective plumbing is what is under test)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_render(params):
    def rast(camera, dims, sh_degree):
        a, b = params
        xys = (a[:, :2] * camera).clone()
        if xys.requires_grad:
            xys.retain_grad()
        img = (xys.sum(-1, keepdim=True) * b).tanh()
        return img, {"xys": xys}
    return rast


def _worker(rank, world, port, overlap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tinysplat_b200.parallel import DataParallelRenderer
    g = torch.Generator().manual_seed(0)
    a = torch.randn(50, 3, generator=g).requires_grad_(True)       # same replica on every rank
    b = torch.randn(50, 4, generator=g).requires_grad_(True)
    dp = DataParallelRenderer(_fake_render((a, b)), [a, b], average=True, overlap=overlap)
    cam = float(rank + 1)                                            # a different "view" per rank
    loss, img, extras = dp.step(cam, None, 0, lambda im, ex: im.pow(2).sum())
    stat = dp.reduce_densify_stat(extras)
    # reference: average of the per-view gradients computed locally
    ga, gb, st = torch.zeros_like(a), torch.zeros_like(b), torch.zeros(50)
    for r in range(world):
        a2 = a.detach().clone().requires_grad_(True)
        b2 = b.detach().clone().requires_grad_(True)
        im, ex = _fake_render((a2, b2))(float(r + 1), None, 0)
        im.pow(2).sum().backward()
        ga += a2.grad / world
        gb += b2.grad / world
        st += ex["xys"].grad.norm(dim=-1)
    ok = torch.allclose(a.grad, ga, atol=1e-6) and torch.allclose(b.grad, gb, atol=1e-6) and \
        torch.allclose(stat, st, atol=1e-6)
    q.put((rank, bool(ok), dp.reducer.payload_bytes()))
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_gradient_allreduce_two_ranks_gloo(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == (50 * 3 + 50 * 4) * 4


def test_single_process_is_a_noop():
    from tinysplat_b200.parallel import GradientAllReducer
    a = torch.randn(4, 3, requires_grad=True)
    red = GradientAllReducer([a])
    (a * 2).sum().backward()
    red.finish()
    assert red.world_size == 1 and torch.allclose(a.grad, torch.full_like(a, 2.0))


# ---- packed-row exchange (PackedGradExchange): the collective plumbing on gloo -----------------
def _packed_worker(rank, world, port, n_gauss, q):
    """Each rank holds 'packed rows' P_r[N, 12] of its own view.  A linear stand-in for the shard
    kernels (sum over views of row * (view + 1)) runs on the rank's shard; after the all-gather
    every rank must hold sum_r P_r * (r + 1) for ALL Gaussians, scaled by 1/world."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tinysplat_b200.parallel import PackedGradExchange
    ex = PackedGradExchange(average=True)
    N = n_gauss
    Ns = ex.shard_rows(N)
    rows = [torch.randn(N, 12, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
    send = torch.zeros(world * Ns, 12)
    send[:N] = rows[rank]
    cams = ex.gather_cameras(torch.full((32,), float(rank)))
    recv = ex.all_to_all_rows(send)
    s0 = rank * Ns
    ns = max(0, min(N, s0 + Ns) - s0)
    shard_a = torch.zeros(Ns, 12)
    shard_b = torch.zeros(Ns, 3)
    for v in range(world):
        shard_a[:ns] += recv[v, :ns] * (cams[v, 0] + 1.0) * ex.out_scale()
        shard_b[:ns] += recv[v, :ns, :3] * ex.out_scale()
    full_a, full_b = ex.all_gather_shards([shard_a, shard_b])
    want_a = sum(rows[r] * (r + 1.0) for r in range(world)) / world
    want_b = sum(rows[r][:, :3] for r in range(world)) / world
    ok = torch.allclose(full_a[:N], want_a, atol=1e-6) and torch.allclose(full_b[:N], want_b, atol=1e-6) \
        and full_a.shape[0] == world * Ns and Ns % 128 == 0 and torch.equal(cams[:, 0], torch.arange(world).float())
    q.put((rank, bool(ok), ex.last_bytes_sent))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_gauss", [1000, 129, 5])     # ragged last shard; shards past the end
def test_packed_exchange_collectives_two_ranks_gloo(n_gauss):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_packed_worker, args=(r, 2, port, n_gauss, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(b > 0 for _, _, b in res)


def test_peer_layout_segments_are_aligned_and_shards_cover_all_rows():
    """Host arithmetic of the peer-memory exchange (parallel.PeerLayout): 256-byte aligned,
    non-overlapping segments; shards of 256-row multiples that cover [0, N) exactly once; gradient
    segments large enough for the capacity; a smaller N fits the same allocation."""
    from tinysplat_b200.parallel import PeerLayout
    for world in (1, 2, 3, 8):
        for cap in (1, 1000, 1_126_024):
            L = PeerLayout(world, cap, 16, 2048)
            spans = sorted(L.seg.values())
            for (o, n), (o2, _) in zip(spans, spans[1:] + [(L.total_bytes, 0)]):
                assert o % 256 == 0 and o + n <= o2
            assert L.seg["geo"][1] == world * (L.cap_shard + L.MAX_CHUNKS * L.SHARD_ALIGN) * 32
            assert L.seg["rgb"][1] == world * world * L.cap_shard * 12
            for name in ("rest", "dc", "means", "scales", "quats", "logit"):
                assert L.seg["g_" + name][1] >= cap * L.width(name) * 4
            for n in (cap, max(1, cap // 2), max(1, cap - 3)):
                covered = 0
                for r in range(world):
                    s0, ns, ns_all = L.shard_of(r, n)
                    assert ns_all % 256 == 0 and ns_all <= L.cap_shard
                    assert s0 == r * ns_all and 0 <= ns <= ns_all
                    if ns:
                        assert s0 == covered
                    covered += ns
                assert covered == n
                # the same rows pushed in pieces (PeerLayout.chunks): contiguous pieces that cover [0, n),
                # every piece sharded on its own in 256-row multiples, the pieces' geometry blocks
                # disjoint and inside the geo segment
                for k in (1, 3, 4, 8, 50):
                    pieces = L.chunks(n, k)
                    assert 1 <= len(pieces) <= min(k, L.MAX_CHUNKS)
                    row, geo_rows = 0, 0
                    for r0, rows, ns_c, g0 in pieces:
                        assert r0 == row and rows > 0 and r0 % (world * 256) == 0
                        assert ns_c % 256 == 0 and ns_c * world >= rows and g0 == geo_rows
                        mine = sum(max(0, min(rows, (r + 1) * ns_c) - r * ns_c) for r in range(world))
                        assert mine == rows
                        row += rows
                        geo_rows += ns_c
                    assert row == n
                    assert world * geo_rows * 32 <= L.seg["geo"][1]
