"""Sort-first multi-GPU plumbing (SURVEY.md §8e — no reference counterpart: the reference is single-GPU).

One process per GPU (torchrun).  Geometry is replicated (broadcast once from rank 0), screen ownership is
interleaved in 64-px super-tiles, ``owner(sx, sy) = (sx + 3*sy) % world`` (ifr_tile_owner in
csrc/common.cuh); each rank bins / rasterises / shades only the tiles it owns, packs its super-tiles
contiguously (pack_kernel) and rank 0 gathers them over NCCL (NVLink) and scatters them into the final
frame (unpack).  There is exactly one exchange step; nothing on the data path before it communicates.

The pure-numpy ``pack_host`` / ``unpack_host`` restate the device layout and are what the gloo CPU tests
exercise (world_size 2) together with ``gather_frame``."""
import numpy as np

SUPER = 64


def owned_supertiles(width, height, rank, world):
    sx_n, sy_n = (width + SUPER - 1) // SUPER, (height + SUPER - 1) // SUPER
    return [(sx, sy) for sy in range(sy_n) for sx in range(sx_n) if (sx + 3 * sy) % world == rank]


def owned_mask(width, height, rank, world):
    ys, xs = np.mgrid[0:height, 0:width]
    return ((xs // SUPER) + 3 * (ys // SUPER)) % world == rank


def pack_host(color, rank, world):
    """color [H, W, 4] float32 -> [n_owned, 64, 64, 4] (edge super-tiles zero padded)."""
    h, w = color.shape[:2]
    tiles = owned_supertiles(w, h, rank, world)
    out = np.zeros((len(tiles), SUPER, SUPER, 4), dtype=np.float32)
    for i, (sx, sy) in enumerate(tiles):
        blk = color[sy * SUPER:(sy + 1) * SUPER, sx * SUPER:(sx + 1) * SUPER]
        out[i, :blk.shape[0], :blk.shape[1]] = blk
    return out


def unpack_host(dst, packed, rank, world):
    h, w = dst.shape[:2]
    packed = np.asarray(packed).reshape(-1, SUPER, SUPER, 4)
    for i, (sx, sy) in enumerate(owned_supertiles(w, h, rank, world)):
        blk = dst[sy * SUPER:(sy + 1) * SUPER, sx * SUPER:(sx + 1) * SUPER]
        blk[...] = packed[i, :blk.shape[0], :blk.shape[1]]
    return dst


def max_owned_floats(width, height, world):
    return max(len(owned_supertiles(width, height, r, world)) for r in range(world)) * SUPER * SUPER * 4


def broadcast_geometry(vb, idx, dist, device):
    """rank 0 holds numpy (vb, idx); every rank returns torch tensors on `device` (ncclBroadcast)."""
    import torch
    rank = dist.get_rank()
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == 0:
        meta = torch.tensor([vb.shape[0], vb.shape[1], idx.shape[0]], dtype=torch.int64, device=device)
    dist.broadcast(meta, 0)
    nv, nf, ni = (int(x) for x in meta.tolist())
    tvb = torch.from_numpy(np.ascontiguousarray(vb, np.float32)).to(device) if rank == 0 else \
        torch.empty((nv, nf), dtype=torch.float32, device=device)
    tib = torch.from_numpy(np.ascontiguousarray(idx, np.int32)).to(device) if rank == 0 else \
        torch.empty(ni, dtype=torch.int32, device=device)
    dist.broadcast(tvb, 0)
    dist.broadcast(tib, 0)
    return tvb, tib


def gather_frame(packed_local, width, height, dist, dst=0):
    """Gather every rank's packed super-tiles on rank `dst`.

    packed_local: 1-D float32 torch tensor (CPU for gloo, CUDA for NCCL) padded to
    max_owned_floats().  Returns the list of per-rank tensors on `dst`, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return [packed_local]
    outs = [torch.empty_like(packed_local) for _ in range(world)] if rank == dst else None
    dist.gather(packed_local, outs, dst=dst)
    return outs


def copy_owned_host(dst, color, rank, world):
    """Host restatement of iftrcuCopyOwnedTilesToHost: the super-tiles `rank` owns of color [H, W, 4] -> the full-frame image dst."""
    m = owned_mask(color.shape[1], color.shape[0], rank, world)
    dst[m] = color[m]
    return dst


class SharedHostImage:
    """One full-frame RGBA32F host image [H, W, 4] mapped by every rank of the box — the target of the distributed read-back
    (iftrcuCopyOwnedTilesToHost): a file under /dev/shm created by rank 0, mapped by all, unlinked once everyone holds the mapping.
    `array` is the numpy view, `ptr` its address (pinned + mapped by the renderer on first use)."""

    def __init__(self, width, height, dist=None):
        import mmap
        import os
        rank = dist.get_rank() if dist is not None else 0
        world = dist.get_world_size() if dist is not None else 1
        self.nbytes = width * height * 16
        name = [None]
        fd = -1
        if rank == 0:
            name = [f"/dev/shm/ifrit_b200_frame_{os.getpid()}_{id(self):x}"]
            fd = os.open(name[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, self.nbytes)
        if world > 1:
            dist.broadcast_object_list(name, 0)
        if rank != 0:
            fd = os.open(name[0], os.O_RDWR)
        self.mm = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        if world > 1:
            dist.barrier()
        if rank == 0:
            os.unlink(name[0])
        self.array = np.frombuffer(self.mm, dtype=np.float32).reshape(height, width, 4)
        self.ptr = self.array.ctypes.data


def setup_slice(ntri, rank, world):
    """Distributed triangle setup: (chunk, first, count) — rank `rank` classifies triangles [first, first + count); chunks are equal
    and 16-aligned so that the all-gather of the owner masks is one contiguous, evenly split tensor."""
    chunk = ((ntri + world - 1) // world + 15) // 16 * 16
    first = rank * chunk
    return chunk, first, max(0, min(ntri, first + chunk) - first)


def box_owner_mask_host(world, tx0, tx1, ty0, ty1):
    """Host restatement of box_owner_mask (csrc/kernels_geometry.cu): bit r set <=> rank r owns a 64-px super-tile that the box of
    16-px tiles [tx0..tx1] x [ty0..ty1] touches; boxes of more than 16 super-tiles name every rank."""
    st = SUPER // 16
    sx0, sx1, sy0, sy1 = tx0 // st, tx1 // st, ty0 // st, ty1 // st
    if (sx1 - sx0 + 1) * (sy1 - sy0 + 1) > 16:
        return (1 << world) - 1
    m = 0
    for sy in range(sy0, sy1 + 1):
        for sx in range(sx0, sx1 + 1):
            m |= 1 << ((sx + 3 * sy) % world)
    return m


def share_framebuffer(renderer, width, height, dist, device, ring=1):
    """Peer-memory sort-first: rank 0 binds a renderer-owned device framebuffer and exports it (CUDA IPC);
    every other rank binds rank 0's memory as its render target, so its tile kernel writes the tiles it owns
    straight into the final frame over NVLink.  Call on every rank before the first draw.  ring = 2: two frames
    back to back in rank 0's allocation (iftrcuSetFrameRing), alternated by SortFirstRenderer.gather()."""
    import torch
    rank = dist.get_rank()
    blob = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        renderer.bindFrameBuffer(None, None, width, height)
        if ring > 1:
            renderer.setFrameRing(ring)
        renderer.synchronize()      # the frame-flag blocks are zeroed before any peer can reach them
        blob = torch.frombuffer(bytearray(renderer.exportFrameBuffer()), dtype=torch.uint8).to(device)
    dist.broadcast(blob, 0)
    if rank != 0:
        renderer.bindPeerFrameBuffer(bytes(blob.cpu().numpy().tobytes()), width, height)
        if ring > 1:
            renderer.setFrameRing(ring)


class SortFirstRenderer:
    """Wraps one TileRasterRendererCuda per rank.  mode "peer" (default on NVLink): every rank renders its
    tiles straight into rank 0's framebuffer (share_framebuffer); frame completion and buffer re-use are ordered
    on the device through flag words in rank 0's memory (iftrcuFrameSignal / Wait / Release / Fence) — no NCCL
    call and no host round trip per frame.  mode "nccl": pack -> NCCL gather -> unpack."""

    def __init__(self, renderer, dist, device, same_stream=True, mode="nccl", clears_every_frame=False, ring=1):
        """same_stream: the renderer was given torch's current stream (setStream), so pack -> NCCL ->
        unpack are ordered on the device and need no host synchronisation."""
        import torch
        self.r, self.dist, self.device, self.same_stream = renderer, dist, device, same_stream
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.mode = mode
        # peer mode: the caller promises that every frame starts with clear() -> rank 0 pre-clears the other ranks' tiles when
        # it releases a frame and the peers keep the colour of their empty tiles off NVLink (iftrcuFrameReleaseCleared)
        self.clears_every_frame = clears_every_frame
        # peer mode, ring = 2: frame k lives in buffer (k - 1) & 1 of rank 0's allocation.  A peer fences frame k + 1 on the release of
        # frame k - 1 (the last user of that buffer) instead of frame k: rank 0's wait / pre-clear / release of frame k overlap the peers'
        # tile kernels of frame k + 1, and the flag round trip release -> fence leaves the critical path.
        self.ring = ring if (mode == "peer" and self.world > 1) else 1
        if mode == "peer" and self.world > 1:
            share_framebuffer(renderer, renderer.width, renderer.height, dist, device, ring=self.ring)
            self.frame = 0
        renderer.setTilePartition(self.rank, self.world)
        self.torch = torch
        self.packed = None
        self._mask = None

    def distribute_setup(self, ntri):
        """Before drawElements(): split the triangle setup over the ranks (peer / nccl mode alike).  Every rank classifies its 1 / world
        slice of the draw's triangles (which ranks own a tile of each), the slices are all-gathered over NCCL on the renderer's stream,
        and the following draw assembles only the triangles that name this rank (iftrcuClassifyTriangles / SetTriangleOwnerMask).  Must
        be called again whenever the geometry or the transform changes — i.e. every frame of an animated scene."""
        torch, r = self.torch, self.r
        if self.world < 2 or self.world > 8:
            return
        chunk, first, count = setup_slice(ntri, self.rank, self.world)
        if self._mask is None or self._mask.numel() != chunk * self.world:
            self._mask = torch.zeros(chunk * self.world, dtype=torch.uint8, device=self.device)
        # everything on the renderer's GEOMETRY stream: classification -> all-gather -> (next draw's) setup pass, so that the whole
        # chain overlaps the previous frame's tile kernel on the draw stream
        r.classifyTriangles(first, count, self._mask.data_ptr())
        g = torch.cuda.ExternalStream(r.geometryStream(), device=self.device)
        with torch.cuda.stream(g):
            self.dist.all_gather_into_tensor(self._mask, self._mask[first:first + chunk])
        r.setTriangleOwnerMask(self._mask.data_ptr())

    def gather(self, hold=False):
        """After drawElements(): rank 0's device framebuffer becomes the complete frame (in stream order on rank 0).
        hold=True (peer mode): rank 0 keeps the frame until release() — e.g. until its read-back is enqueued."""
        torch, r = self.torch, self.r
        if self.world == 1:
            return
        if self.mode == "peer":
            self.frame = (self.frame + 1) & 0xffffffff   # compared modulo 2^32 on the device
            r.frameSignal(self.frame)             # my tiles of this frame are stored (after my tile kernel)
            if self.rank == 0:
                if not hold and self.ring > 1:
                    # nothing on the draw stream consumes the frame: wait / pre-clear / release run on the exchange stream and
                    # overlap this rank's tile kernel of the next frame (other buffer of the ring)
                    r.frameRetire(self.frame, cleared=self.clears_every_frame)
                    r.selectFrame(self.frame & 1)
                else:
                    r.frameWait(self.frame)           # polls until every rank has signalled
                    if not hold:
                        self.release()
            elif self.ring == 1:
                r.frameFence(self.frame, cleared=self.clears_every_frame)   # my next tile kernel waits until rank 0 has released this frame
            else:
                r.selectFrame(self.frame & 1)     # buffer of frame k + 1, last used by frame k - 1
                if self.frame >= 2:
                    r.frameFence(self.frame - 1, cleared=self.clears_every_frame)
            return
        n = max_owned_floats(r.width, r.height, self.world)
        if self.packed is None or self.packed.numel() != n:
            self.packed = torch.zeros(n, dtype=torch.float32, device=self.device)
        r.packOwnedTiles(self.packed.data_ptr(), n * 4)
        if not self.same_stream:
            r.synchronize()   # pack ran on the renderer's own stream; NCCL is ordered after torch's
        outs = gather_frame(self.packed, r.width, r.height, self.dist)
        if self.rank == 0:
            if not self.same_stream:
                torch.cuda.current_stream().synchronize()
            for src in range(1, self.world):
                r.unpackOwnedTiles(outs[src].data_ptr(), src, self.world)
            if not self.same_stream:
                r.synchronize()
            self._keep = outs   # keep the receive buffers alive until the unpack kernels have run

    def read_back_distributed(self, image):
        """After drawElements() into a LOCAL framebuffer: this rank's tiles go straight to the shared host image (SharedHostImage) over
        its own PCIe link — no NVLink gather, N links instead of one.  The frame is complete once every rank's stream has passed."""
        self.r.copyOwnedTilesToHost(image.ptr)

    def release(self):
        """Peer mode, after gather(hold=True): rank 0 lets the other ranks overwrite the frame."""
        if self.world > 1 and self.mode == "peer" and self.rank == 0:
            self.r.frameRelease(self.frame, cleared=self.clears_every_frame)
            if self.ring > 1:
                self.r.selectFrame(self.frame & 1)
