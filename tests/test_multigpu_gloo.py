"""CPU test of the N>1 path (world_size 2, gloo): sort-first partition + gather plumbing.

Each rank "renders" only the super-tiles it owns (here: the oracle's frame masked to the rank's tiles,
because the CUDA path cannot run without a GPU), packs them, rank 0 gathers with torch.distributed and
unpacks; the result must be bit-identical to the single-rank frame."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, W, H, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ifrit_v2_b200  # noqa: F401
    from ifrit_v2_b200 import scenes, multigpu as MG
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # geometry lives on rank 0 and is broadcast (replicated geometry, sort-first)
    vb = idx = None
    if rank == 0:
        vb, idx = scenes.mesh_sphere(n=24)
    tvb, tib = MG.broadcast_geometry(vb, idx, dist, "cpu")
    vb, idx = tvb.numpy(), tib.numpy()
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    st = O.make_state(W, H, msaa=4, cull=O.CULL_BACK, clear_color=(.2, .3, .4, 1))
    full = O.draw(st, vb, idx, O.VS_MVP, O.FS_NORMAL_COLOR, 2, u)["color"]
    mine = np.where(MG.owned_mask(W, H, rank, world)[..., None], full, np.float32(-1))   # only owned tiles are valid
    n = MG.max_owned_floats(W, H, world)
    packed = torch.zeros(n, dtype=torch.float32)
    p = MG.pack_host(mine, rank, world).reshape(-1)
    packed[: p.size] = torch.from_numpy(p)
    outs = MG.gather_frame(packed, W, H, dist)
    if rank == 0:
        frame = np.full_like(full, -2)
        for src in range(world):
            MG.unpack_host(frame, outs[src].numpy()[: len(MG.owned_supertiles(W, H, src, world)) * 64 * 64 * 4], src, world)
        q.put(bool(np.array_equal(frame, full)))
    dist.barrier()
    dist.destroy_process_group()


def _worker_shared(rank, world, port, W, H, q):
    import sys
    sys.path.insert(0, ROOT)
    import ifrit_v2_b200  # noqa: F401
    from ifrit_v2_b200 import multigpu as MG
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.random.default_rng(7).random((H, W, 4), dtype=np.float32)       # the same "frame" on every rank
    img = MG.SharedHostImage(W, H, dist)
    if rank == 0:
        img.array[...] = -1.0
    dist.barrier()
    MG.copy_owned_host(img.array, full, rank, world)       # what iftrcuCopyOwnedTilesToHost does on the device side
    dist.barrier()
    if rank == 0:
        q.put(bool(np.array_equal(img.array, full)))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_read_back_shared_image_gloo():
    """Sort-first distributed read-back: every rank stores the tiles it owns into ONE host image mapped by all ranks
    (multigpu.SharedHostImage); the union is the full frame (world_size 2, odd frame size)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_shared, args=(r, 2, port, 400, 230, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


@pytest.mark.parametrize("world", [2])
def test_sort_first_gather_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, 400, 230, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_peer_mode_frame_flag_protocol():
    """SortFirstRenderer in peer mode drives the device-side frame flags (include/ifrit_b200.h: iftrcuFrameSignal / Wait /
    Release / Fence): every rank signals its frame; rank 0 waits for all and releases (or holds until release()); the
    other ranks fence their next tile kernel on that release.  Host logic only, with a recording stand-in renderer."""
    from ifrit_v2_b200 import multigpu as MG

    class Rec:
        width, height = 128, 64

        def __init__(self):
            self.calls = []

        def setTilePartition(self, rank, world):
            self.calls.append(("partition", rank, world))

        def __getattr__(self, name):
            if name.startswith("frame"):
                return lambda fid, cleared=False: self.calls.append((name, fid))
            raise AttributeError(name)

    class Dist:
        def __init__(self, rank, world):
            self.rank, self.world = rank, world

        def get_rank(self):
            return self.rank

        def get_world_size(self):
            return self.world

    share = MG.share_framebuffer
    MG.share_framebuffer = lambda *a, **k: None          # CUDA IPC export / bind: exercised by tests/mgpu_check.py
    try:
        r0, r1 = Rec(), Rec()
        s0 = MG.SortFirstRenderer(r0, Dist(0, 2), None, mode="peer")
        s1 = MG.SortFirstRenderer(r1, Dist(1, 2), None, mode="peer")
        s0.gather(); s1.gather()
        s0.gather(hold=True); s1.gather(hold=True)
        s0.release(); s1.release()
    finally:
        MG.share_framebuffer = share
    assert r0.calls == [("partition", 0, 2), ("frameSignal", 1), ("frameWait", 1), ("frameRelease", 1),
                        ("frameSignal", 2), ("frameWait", 2), ("frameRelease", 2)]
    assert r1.calls == [("partition", 1, 2), ("frameSignal", 1), ("frameFence", 1), ("frameSignal", 2), ("frameFence", 2)]


def test_peer_mode_frame_ring_protocol():
    """Two frames in flight (iftrcuSetFrameRing / iftrcuSelectFrame): frame k lives in buffer (k - 1) & 1; rank 0 releases frame k and
    moves on to the other buffer; a peer moves on first and fences its next tile kernel on the release of frame k - 1 — the last user of
    that buffer — and on nothing at all before frame 2.  Host logic only, with a recording stand-in renderer."""
    from ifrit_v2_b200 import multigpu as MG

    class Rec:
        width, height = 128, 64

        def __init__(self):
            self.calls = []

        def setTilePartition(self, rank, world):
            pass

        def selectFrame(self, index):
            self.calls.append(("select", index))

        def __getattr__(self, name):
            if name.startswith("frame"):
                return lambda fid, cleared=False: self.calls.append((name, fid, cleared))
            raise AttributeError(name)

    class Dist:
        def __init__(self, rank, world):
            self.rank, self.world = rank, world

        def get_rank(self):
            return self.rank

        def get_world_size(self):
            return self.world

    share = MG.share_framebuffer
    seen = []
    MG.share_framebuffer = lambda *a, **k: seen.append(k.get("ring"))
    try:
        r0, r1 = Rec(), Rec()
        s0 = MG.SortFirstRenderer(r0, Dist(0, 2), None, mode="peer", clears_every_frame=True, ring=2)
        s1 = MG.SortFirstRenderer(r1, Dist(1, 2), None, mode="peer", clears_every_frame=True, ring=2)
        for _ in range(3):
            s0.gather(); s1.gather()
        s0.gather(hold=True); s1.gather(hold=True)
        s0.release(); s1.release()
    finally:
        MG.share_framebuffer = share
    assert seen == [2, 2]
    # frames nothing consumes are retired on the exchange stream (wait + pre-clear + release in one verb); a held frame is waited
    # for on the draw stream and released by release()
    assert r0.calls == [("frameSignal", 1, False), ("frameRetire", 1, True), ("select", 1),
                        ("frameSignal", 2, False), ("frameRetire", 2, True), ("select", 0),
                        ("frameSignal", 3, False), ("frameRetire", 3, True), ("select", 1),
                        ("frameSignal", 4, False), ("frameWait", 4, False), ("frameRelease", 4, True), ("select", 0)]
    assert r1.calls == [("frameSignal", 1, False), ("select", 1),
                        ("frameSignal", 2, False), ("select", 0), ("frameFence", 1, True),
                        ("frameSignal", 3, False), ("select", 1), ("frameFence", 2, True),
                        ("frameSignal", 4, False), ("select", 0), ("frameFence", 3, True)]


def _worker_masks(rank, world, port, ntri, q):
    import sys
    sys.path.insert(0, ROOT)
    import ifrit_v2_b200  # noqa: F401
    from ifrit_v2_b200 import multigpu as MG
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)                      # the same boxes on every rank (replicated geometry)
    tx0 = rng.integers(0, 230, ntri); ty0 = rng.integers(0, 130, ntri)
    tx1 = tx0 + rng.integers(0, 6, ntri) * (rng.random(ntri) < 0.3); ty1 = ty0 + rng.integers(0, 6, ntri) * (rng.random(ntri) < 0.3)
    tx1[:3] = tx0[:3] + 40; ty1[:3] = ty0[:3] + 40      # a few boxes of more than 16 super-tiles
    chunk, first, count = MG.setup_slice(ntri, rank, world)
    mask = torch.zeros(chunk * world, dtype=torch.uint8)
    for t in range(first, first + count):                # what iftrcuClassifyTriangles writes for this rank's slice
        mask[t] = MG.box_owner_mask_host(world, int(tx0[t]), int(tx1[t]), int(ty0[t]), int(ty1[t]))
    dist.all_gather_into_tensor(mask, mask[first:first + chunk].clone())
    ok = True
    if rank == 0:
        m = mask.numpy()
        # every slice arrived where it belongs, and the masks name (at least) every rank that owns a tile of the box
        for t in range(ntri):
            want = 0
            for ty in range(int(ty0[t]), int(ty1[t]) + 1):
                for tx in range(int(tx0[t]), int(tx1[t]) + 1):
                    want |= 1 << (((tx // 4) + 3 * (ty // 4)) % world)
            ok = ok and (int(m[t]) & want) == want and int(m[t]) == MG.box_owner_mask_host(world, int(tx0[t]), int(tx1[t]), int(ty0[t]), int(ty1[t]))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_setup_slices_and_owner_masks_gloo():
    """Distributed triangle setup, host side (world_size 2): the slices tile the draw (16-aligned, equal chunks, ragged tail), each rank
    classifies its own, one all-gather gives everyone the whole mask array, and a mask names every rank that owns a tile of the box."""
    from ifrit_v2_b200 import multigpu as MG
    for ntri, world in ((1, 2), (31, 2), (1000, 3), (50_000_000, 8)):
        seen = 0
        for rank in range(world):
            chunk, first, count = MG.setup_slice(ntri, rank, world)
            assert chunk % 16 == 0 and chunk * world >= ntri and first == rank * chunk and first + count <= max(ntri, first)
            seen += count
        assert seen == ntri
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_masks, args=(r, 2, port, 1003, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
