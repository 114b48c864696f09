#!/usr/bin/env python
"""bench.py — the measurement contract.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2|c3|c4|c5]

One "step" = one frame of the hot path: clear + drawElements (+ the sort-first gather when N > 1) of the
named BASELINE.json configuration on synthetic, seeded geometry (ifrit-v2_b200/scenes.py).  Prints ONE
JSON line (rank 0).  See DESIGN.md §6 for the definition of every field.

  value         whole-job Mtris/s (triangles submitted to drawElements / frame time), inputs resident
                in HBM, CUDA-event timed on the launching stream, max over ranks.
  e2e           same metric through the C-ABI with HOST buffers: per step the vertex + index buffers are
                bound from pinned host memory (H2D), the frame is drawn and the colour image is copied
                back to the host image (D2H), all inside the timed region.
  roofline      dominant kernel (tile kernel): algorithmic bytes per launch / its CUDA-event duration
                against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the UNMODIFIED reference CPU renderer (oracle/_ref/libifrit_ref.so, 17 threads) timed on
                this box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1..4]; scene definitions in SURVEY.md §8(d)
    "c2": dict(workload="mesh50k 3840x2160 textured (2048^2 RGBA32F, 11 mips, aniso16) + depth test, 1x",
               W=3840, H=2160, scene="mesh", msaa=1, cull=0, fs="textured"),
    "c3": dict(workload="mesh50k 3840x2160 8xMSAA textured aniso16, back-face + contribution cull, ddx/ddy",
               W=3840, H=2160, scene="mesh", msaa=8, cull=2, fs="textured"),
    "c4": dict(workload="soup10M 3840x2160 small-triangle soup (binning/atomic stress), 1x",
               W=3840, H=2160, scene="soup", ntri=10_000_000, msaa=1, cull=2, fs="varying0"),
    "c5": dict(workload="scene50M 7680x4320 8xMSAA, sort-first over N GPUs",
               W=7680, H=4320, scene="mix", ntri=50_000_000, msaa=8, cull=2, fs="varying0"),
}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                mhz.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz), "window": "untimed pre-roll of the same steps + the timed region, 20 ms period"}


def build_scene(cfg, small=False):
    import ifrit_v2_b200  # noqa: F401
    from ifrit_v2_b200 import scenes
    W, H = cfg["W"], cfg["H"]
    if cfg["scene"] == "mesh":
        vb, idx = scenes.mesh_sphere(n=160, seed=1234)
        u = scenes.uniform_block(scenes.default_mvp(W, H))
        tex = scenes.texture_checker(size=2048 if not small else 256, seed=7)
        return dict(vb=vb, idx=idx, u=u, tex=tex, vs=1, nvary=2, mips=11 if not small else 8)
    import numpy as np
    gen = scenes.soup if cfg["scene"] == "soup" else scenes.scene_mix
    vb, idx = gen(ntri=cfg["ntri"], width=W, height=H)
    return dict(vb=vb, idx=idx, u=scenes.uniform_block(np.eye(4)), tex=None, vs=0, nvary=1, mips=0)


def algorithmic_bytes(cfg, sc, covered_pixels):
    """B = V*Sv + I*4 + P*(16+4) + Tex, Tex = min(1.33*covered, texels in chain)*16."""
    V, Sv = sc["vb"].shape[0], sc["vb"].shape[1] * 4
    I = sc["idx"].shape[0]
    P = cfg["W"] * cfg["H"]
    tex = 0
    if sc["tex"] is not None:
        n, w = 0, sc["tex"].shape[0]
        for _ in range(sc["mips"] + 1):
            n += w * w
            w = (w + 1) >> 1
        tex = int(min(1.33 * covered_pixels, n)) * 16
    geometry = V * Sv + I * 4
    tile = P * 20 + tex
    return dict(frame=geometry + tile, geometry=geometry, tile=tile, tex=tex)


# ------------------------------------------------------------------------------------------------------
def run_ours(args, cfg_name, do_e2e=True, fast_sampler=True):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ifrit_v2_b200  # noqa: F401
    from ifrit_v2_b200 import renderer as R, multigpu as MG

    cfg = CONFIGS[cfg_name]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    W, H = cfg["W"], cfg["H"]

    sc = build_scene(cfg) if rank == 0 or world == 1 else None
    if world > 1:
        # geometry is replicated: broadcast once from rank 0 (static scene)
        import numpy as _np
        meta = [None]
        if rank == 0:
            meta = [dict(u=sc["u"], tex=sc["tex"], vs=sc["vs"], nvary=sc["nvary"], mips=sc["mips"])]
        dist.broadcast_object_list(meta, 0)
        tvb, tib = MG.broadcast_geometry(sc["vb"] if rank == 0 else None, sc["idx"] if rank == 0 else None, dist, dev)
        if rank != 0:
            sc = dict(meta[0])
            sc["vb"], sc["idx"] = _np.empty((tvb.shape[0], tvb.shape[1]), _np.float32), _np.empty(tib.shape[0], _np.int32)
    else:
        tvb = torch.from_numpy(sc["vb"]).to(dev)
        tib = torch.from_numpy(sc["idx"]).to(dev)
    ntri = tib.shape[0] // 3

    r = R.TileRasterRendererCuda(local)
    r.Init()
    stream = torch.cuda.Stream(device=dev)   # one explicit stream for kernels, NCCL ordering and the timing events
    torch.cuda.set_stream(stream)
    r.setStream(stream.cuda_stream)
    mg_mode = args.mgpu if world > 1 else None
    color = depth = None
    if mg_mode == "peer":
        r.width, r.height = W, H            # framebuffer: rank 0 owns it, the others bind it over NVLink (CUDA IPC)
    else:
        color = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        depth = torch.zeros((H, W), dtype=torch.float32, device=dev)
        r.bindFrameBufferDevice(color, depth, W, H)

    def configure(rr):
        rr.bindVertexBuffer(tvb, tvb.shape[0], tvb.shape[1] // 4, device=True)
        rr.bindIndexBuffer(tib, tib.shape[0], device=True)
        rr.bindVertexShader(sc["vs"], sc["nvary"])
        rr.bindFragmentShader(cfg["fs"])
        rr.CreateBuffer(0, sc["u"].nbytes)
        rr.copyHostBufferToBuffer(sc["u"], 0)
        if sc["tex"] is not None:
            s = sc["tex"].shape[0]
            rr.CreateTexture(0, s, s, sc["mips"], 1)
            rr.copyHostBufferToImage(np.ascontiguousarray(sc["tex"][None]), 0, s, s)
            rr.generateMipmap(0, R.IF_FILTER_LINEAR)
            rr.CreateSampler(0, filterMode=R.IF_FILTER_LINEAR, anisotropyEnable=True, maxAnisotropy=16.0)
        rr.setMsaaSamples(cfg["msaa"])
        rr.SetCullMode(cfg["cull"])
        rr.SetDepthFunc(R.IF_COMPARE_OP_LESS)
        rr.SetDepthTestEnable(True)
        rr.setClearValues([(0, 0, 0, 0)], 255.0)   # reference demo clear depth, SG/demo/DefaultDemo.cpp:355
        if getattr(args, "tile_chunk", 0):
            rr.setTileChunk(args.tile_chunk)     # A/B: pin the staged chunk of the tile kernel
        if getattr(args, "no_grid_hint", False):
            rr.setTileGridHint(False)        # A/B: one CTA per tile, the surplus exits
        if getattr(args, "no_tma", False):
            rr.setTmaFramebuffer(False)      # A/B: per-thread framebuffer stores instead of cp.async.bulk.tensor
        rr.setSamplerHardwareFilter(fast_sampler)   # library default: the oracle's arithmetic (software fp32 lerps); the bench opts in to FAST (texture unit)

    configure(r)
    if getattr(args, "emulate_rank", None) and world == 1:
        # analysis aid (1 GPU): render only the tiles rank R of N would own — the per-rank kernel times of an N-GPU run without
        # the exchange; the line is NOT a bench value (value / roofline refer to the whole frame)
        er, en = (int(x) for x in args.emulate_rank.split("/"))
        r.setTilePartition(er, en)
    sf = MG.SortFirstRenderer(r, dist, dev, mode=mg_mode, clears_every_frame=not getattr(args, 'no_preclear', False), ring=getattr(args, "ring", 2)) if world > 1 else None   # every bench frame is clear() + drawElements()

    # N >= 4: draws of a million triangles and more split the triangle setup over the ranks (one all-gather of owner masks per
    # frame, multigpu.SortFirstRenderer.distribute_setup); below that the replicated setup is cheaper than the collective.  The
    # setup pass is memory bound (indices + vertex gather in, records out): with the masks a rank reads 2 / N of the geometry
    # instead of all of it, which is nothing at N = 2 (measured: c5 12.48 vs 11.91 ms, profiles/r02_mgpu.txt) and 3x less at N = 8
    dsetup = bool(sf) and 4 <= world <= 8 and ntri >= (1 << 20) and not getattr(args, "no_distributed_setup", False)
    if getattr(args, "force_distributed_setup", False) and sf and world <= 8:
        dsetup = True

    emu_mask = None
    if getattr(args, "emulate_rank", None) and world == 1 and getattr(args, "force_distributed_setup", False):
        # emulated distributed setup: the whole owner-mask array is classified once up front (stands for the all-gathered slices
        # of the other ranks); every frame re-classifies this rank's 1 / N slice, as distribute_setup would
        er, en = (int(x) for x in args.emulate_rank.split("/"))
        chunk = ((ntri + en - 1) // en + 15) // 16 * 16
        emu_mask = torch.zeros(chunk * en, dtype=torch.uint8, device=dev)
        r.classifyTriangles(0, ntri, emu_mask.data_ptr())
        r.synchronize()
        r.setTriangleOwnerMask(emu_mask.data_ptr())
        emu_first, emu_count = er * chunk, max(0, min(ntri, er * chunk + chunk) - er * chunk)

    def step():
        r.clear()
        if emu_mask is not None:
            r.classifyTriangles(emu_first, emu_count, emu_mask.data_ptr())
        if dsetup:
            sf.distribute_setup(ntri)
        r.drawElements()
        if sf:
            sf.gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    r.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # untimed pre-roll (same steps, about 0.3 s, same count on every rank): clocks settle under load and
    # nvidia-smi gets samples even when the K timed steps last only a few milliseconds; the sampler keeps
    # running through the timed region
    t_pre = time.perf_counter()
    for _ in range(5):
        step()
    r.synchronize()
    est = torch.tensor([(time.perf_counter() - t_pre) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    pre = int(min(max(0.3 / max(float(est.item()), 1e-6), 5), 3000))
    for i in range(pre):
        step()
        if i % 50 == 49:
            r.synchronize()
    r.synchronize()
    st0 = r.stats()
    r.setProfiling(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    ktimes, ndraws = r.kernelTimes()
    r.setProfiling(False)
    st1 = r.stats()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        frags_t = torch.tensor([st1["fragments_shaded"]], dtype=torch.float64, device=dev)
        dist.all_reduce(frags_t)
        frags = float(frags_t.item())
    else:
        frags = float(st1["fragments_shaded"])
    ms_step = ms / args.steps
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    tile_ms_ranks = None
    if world > 1:       # the tile kernel of every rank (rank 0 stores locally, the others over NVLink)
        tt = torch.zeros(world, dtype=torch.float64, device=dev)
        tt[rank] = float(ktimes.get("tile", 0.0))
        dist.all_reduce(tt)
        tile_ms_ranks = [round(float(x), 4) for x in tt.tolist()]

    # single-frame latency: the same steps with frame pipelining off (geometry and tile stage of a frame back to back on
    # one stream, nothing of the next frame overlapping) — `value` is steady-state throughput with it on
    lat_ms, ktimes_alone = None, None
    if world == 1:
        r.setFramePipelining(False)
        for _ in range(3):
            step()
        r.synchronize()
        nl = max(5, min(args.steps, 100))
        r.setProfiling(True)
        e0.record(stream)
        for _ in range(nl):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        lat_ms = e0.elapsed_time(e1) / nl
        ktimes_alone, _ = r.kernelTimes()      # every kernel with the GPU to itself
        r.setProfiling(False)
        r.setFramePipelining(True)

    # ---- N > 1: the gathered frame on rank 0 must equal the single-GPU frame bit for bit --------------------------
    frame_same = None
    ref_frame = None
    if world > 1:
        r.clear()
        r.drawElements()
        sf.gather(hold=True)         # held until it has been read: a (pre-clearing) release lets the frame be overwritten
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            got = r.readColor()
            sf.release()
            one = R.TileRasterRendererCuda(local)
            one.Init()
            one.bindFrameBuffer(None, None, W, H)
            configure(one)
            one.clear()
            one.drawElements()
            one.synchronize()
            ref_frame = one.readColor()
            frame_same = bool(np.array_equal(got.view(np.uint32), ref_frame.view(np.uint32)))
            one.close()
            del got
        dist.barrier()

    # ---- e2e: host buffers through the C-ABI, H2D + D2H inside the timed region -----------------------
    e2e_ms = e2e8_ms = None
    e2e_steps, readback, h2d_bytes = 0, None, 0
    if do_e2e:
        hvb = torch.from_numpy(sc["vb"]).pin_memory() if rank == 0 or world == 1 else torch.empty(tvb.shape, dtype=torch.float32).pin_memory()
        hib = torch.from_numpy(sc["idx"]).pin_memory() if rank == 0 or world == 1 else torch.empty(tib.shape, dtype=torch.int32).pin_memory()
        if world > 1 and rank != 0:
            hvb.copy_(tvb.cpu()); hib.copy_(tib.cpu())
        hcol = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
        e2e_steps = max(3, min(args.steps, 20))

        # N = 1: the reference-facing call sequence — the host image is bound with useDoubleBuffer (the reference's
        # default, TileRasterRendererCuda.h:88), so drawElements() itself copies a frame back on every call: frame k - 1
        # travels to the host image on a second stream while frame k renders (CU:3899-3909).  N > 1: explicit read-back
        # of the gathered frame on rank 0.
        readback = "explicit copy of the gathered frame after the draw (rank 0)"
        if world == 1:
            r.bindFrameBuffer(hcol.numpy(), None, W, H, useDoubleBuffer=True)
            readback = ("double-buffered, reference useDoubleBuffer semantics: every drawElements() copies one full RGBA32F frame "
                        "(frame k-1) to the pinned host image on a second stream while frame k renders")

        def e2e_step():
            r.bindVertexBuffer(hvb, hvb.shape[0], hvb.shape[1] // 4)        # H2D (pinned, async)
            r.bindIndexBuffer(hib, hib.shape[0])                           # H2D
            r.clear()
            r.drawElements()
            if world == 1:
                r.synchronize()                                            # render stream + copy stream
                return
            sf.gather(hold=True)
            if rank == 0:
                r.copyColorToHostAsync(hcol.data_ptr())                    # D2H of the step's result (pinned, async)
            sf.release()         # peer mode: only now may the other ranks overwrite the frame (device-side flag)
            torch.cuda.current_stream().synchronize()

        def time_e2e():
            for _ in range(2):
                e2e_step()
            barrier()
            e0.record(stream)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            e1.record(stream)
            barrier()
            return max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)

        e2e_ms = time_e2e()
        # N > 1, distributed read-back (the e2e headline when it runs): every rank renders its tiles into a LOCAL framebuffer and stores
        # them straight into ONE host image mapped by all ranks (multigpu.SharedHostImage, iftrcuCopyOwnedTilesToHost) — no NVLink
        # gather, N PCIe links instead of rank 0's one.  The gathered variant above stays in the line as e2e_gather_rank0.
        e2e_gather_ms = e2e_dist_same = None
        if world > 1 and not getattr(args, "no_distributed_readback", False):
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            shared = None
            try:
                shared = MG.SharedHostImage(W, H, dist)
            except Exception as ex:      # collective inside: every rank fails or none does
                print(f"[bench] rank {rank}: shared host image unavailable: {ex}", file=sys.stderr)
            if shared is not None:
                try:
                    r.bindFrameBuffer(None, None, W, H)          # local device framebuffer on every rank; the tile partition stays
                    r.clear(); r.drawElements(); r.copyOwnedTilesToHost(shared.ptr); r.synchronize()
                except Exception as ex:
                    print(f"[bench] rank {rank}: distributed read-back unavailable: {ex}", file=sys.stderr)
                    ok.zero_()
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()):
                    def e2e_step():      # noqa: F811 — the step time_e2e() runs from here on
                        r.bindVertexBuffer(hvb, hvb.shape[0], hvb.shape[1] // 4)        # H2D (pinned, async), every rank
                        r.bindIndexBuffer(hib, hib.shape[0])
                        r.clear()
                        r.drawElements()
                        r.copyOwnedTilesToHost(shared.ptr)                             # D2H: this rank's tiles, its own PCIe link
                        torch.cuda.current_stream().synchronize()
                    e2e_gather_ms = e2e_ms
                    if rank == 0:
                        shared.array[...] = -7.0
                    barrier()
                    e2e_ms = time_e2e()
                    if rank == 0 and ref_frame is not None:
                        e2e_dist_same = bool(np.array_equal(shared.array.view(np.uint32), ref_frame.view(np.uint32)))
                    readback = ("distributed: every rank stores the tiles it owns from its local framebuffer straight into one shared pinned "
                                "host image (iftrcuCopyOwnedTilesToHost), N PCIe links, no NVLink gather; a step ends when every rank's stream has drained")
        # the same call sequence with the packed RGBA8 presentation image bound: the per-draw
        # copy-back carries 4 bytes per pixel instead of 16.  Reported next to `e2e`, never instead of it.
        e2e8_ms = None
        if world == 1:
            hcol8 = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
            r.bindPresentBuffer(hcol8.numpy())
            e2e8_ms = time_e2e() / e2e_steps
            r.bindPresentBuffer(None, 0)
        if world > 1:
            t = torch.tensor([e2e_ms, e2e_gather_ms or 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t[0].item())
            e2e_gather_ms = float(t[1].item()) / e2e_steps if e2e_gather_ms else None
        e2e_ms /= e2e_steps
        h2d_bytes = int(hvb.numel() * 4 + hib.numel() * 4)
        del hvb, hib, hcol

    out = None
    if rank == 0:
        peak, peak_src = peak_hbm()
        covered = frags   # FS invocations of one frame (= covered pixels at 1x, >= covered pixels under MSAA)
        ab = algorithmic_bytes(cfg, sc, covered_pixels=min(covered, W * H))
        tile_ms = ktimes["tile"]
        achieved = ab["tile"] / world / (tile_ms * 1e-3) / 1e9 if tile_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and world == 1:   # dram bytes of one launch from the committed ncu --set full capture (1 GPU only)
            try:
                traffic = json.load(open(tp)).get(cfg_name)
            except Exception:
                traffic = None
        out = {
            "metric": "Mtris/s", "value": ntri / (ms_step * 1e-3) / 1e6, "unit": "Mtris/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "int64 edge functions (24.8 fixed point) + f32 shading", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": cfg_name, "triangles": ntri, "resolution": [W, H],
                       "msaa": cfg["msaa"], "l2": "no explicit flush: per-step working set (framebuffer "
                       f"{W * H * 20 / 1e6:.0f} MB + texture) exceeds the 126 MB L2",
                       "sampler": ("FAST (opt-in): texture-unit bilinear blend, 8-bit weights, SFU lod math" if fast_sampler else
                                   "default: software fp32 lerps, the oracle's arithmetic op for op") if sc["tex"] is not None else "none",
                       "parallelism": (f"sort-first x{world} ({mg_mode}" + (", two frames in flight" if (mg_mode == "peer" and getattr(args, "ring", 2) == 2) else "") + (", pre-cleared release" if not getattr(args, "no_preclear", False) else "") +
                                       (", triangle setup split over the ranks + NCCL all-gather of owner masks every frame" if dsetup else "") + ")") if world > 1 else "1 GPU"},
            "fps": 1e3 / ms_step, "gfrags_per_s": frags / (ms_step * 1e-3) / 1e9,
            "frame_latency_ms": lat_ms,
            "pipelining": "geometry stage of frame k+1 (own stream, second scratch set) overlaps the tile kernel of frame k; frame_latency_ms = same steps with it off",
            "fragments_per_frame": frags,
            "gpu_launches": int(launches),
            "kernel_ms": ktimes, "kernel_ms_draws": ndraws, "kernel_ms_unpipelined": ktimes_alone,
            "roofline": {"bound": "hbm", "kernel": "tile_kernel", "achieved": achieved, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_launch": ab["tile"] / world, "traffic": traffic,
                         "note": "duration from CUDA events around the kernel inside the timed (pipelined) region, where it shares the GPU "
                                 "with the next frame's geometry kernels; frac_unpipelined = the same kernel with the GPU to itself",
                         "frac_unpipelined": (ab["tile"] / world / (ktimes_alone["tile"] * 1e-3) / 1e9 / peak) if ktimes_alone and ktimes_alone["tile"] > 0 else None,
                         # the same fraction from the DRAM bytes ncu measured for one launch (profiles/traffic.json) instead of the algorithmic
                         # ones: SURVEY 8d's texel term saturates at the whole mip chain, the kernel actually touches less of it
                         "frac_measured_traffic": (traffic / (tile_ms * 1e-3) / 1e9 / peak) if traffic and tile_ms > 0 else None},
            "frame_roofline": {"algorithmic_bytes": ab["frame"], "achieved": ab["frame"] / (ms_step * 1e-3) / 1e9,
                               "frac": ab["frame"] / (ms_step * 1e-3) / 1e9 / peak},
            "clocks": clocks,
        }
        if e2e_ms:
            out["e2e"] = {"value": ntri / (e2e_ms * 1e-3) / 1e6, "unit": "Mtris/s", "ms_per_step": e2e_ms,
                          "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(W * H * 16), "steps": e2e_steps,
                          "readback": readback}
        if world > 1:
            out["frame_hash_equals_1gpu"] = frame_same
            out["tile_ms_per_rank"] = tile_ms_ranks
            if e2e_ms and e2e_gather_ms:
                out["e2e"]["frame_equals_1gpu"] = e2e_dist_same
                out["e2e_gather_rank0"] = {"value": ntri / (e2e_gather_ms * 1e-3) / 1e6, "unit": "Mtris/s", "ms_per_step": e2e_gather_ms,
                                           "note": "same steps with the frame gathered on rank 0 over NVLink and copied back over rank 0's PCIe link"}
        if e2e8_ms:
            out["e2e_rgba8"] = {"value": ntri / (e2e8_ms * 1e-3) / 1e6, "unit": "Mtris/s", "ms_per_step": e2e8_ms,
                                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(W * H * 4),
                                "note": "same steps with iftrcuBindPresentBuffer(RGBA8): packed presentation image copied back instead of RGBA32F"}
    r.close()
    del tvb, tib
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------------
def run_reference(args, cfg_name, steps, warmup):
    """The reference's own CPU renderer on this box's host cores, same scene / resolution.  The CPU
    renderer has no sampler library, MSAA or cull-mode state: textured configs use the
    bilinear LOD-0 C fragment shader, BACK culling is what it always does."""
    import numpy as np
    import ifrit_v2_b200  # noqa: F401
    from oracle import oracle as O, ref_driver as RD
    cfg = dict(CONFIGS[cfg_name])
    if not RD.available():
        return None
    sc = build_scene(cfg)
    note = ""
    if cfg["scene"] != "mesh":
        # bounded sample: the vertex upload of the reference C ABI is one call per attribute value
        n = 200_000
        sc["vb"], sc["idx"] = sc["vb"][: 3 * n], sc["idx"][: 3 * n]
        note = f" (first {n} triangles)"
    fs = O.FS_TEXTURED_LOD0 if sc["tex"] is not None else O.FS_VARYING0
    ref = RD.RefRenderer(cfg["W"], cfg["H"], sc["vb"], sc["idx"], sc["vs"], fs, sc["nvary"], sc["u"], texture=sc["tex"])
    for _ in range(warmup):
        ref.draw()
    ts = ref.timed_draws(steps)
    ntri = len(sc["idx"]) // 3
    covered = int((ref.depth() < 1e38).sum())
    med = sorted(ts)[len(ts) // 2]
    mean = sum(ts) / len(ts)
    return dict(ms_per_step=mean * 1e3, median_ms=med * 1e3, min_ms=min(ts) * 1e3, ntri=ntri, covered=covered,
                value=ntri / mean / 1e6, threads=RD.REF_THREADS, cores=os.cpu_count(),
                sample=f"{steps} frames of {cfg['workload']}{note}; CPU path: bilinear LOD-0 FS, no MSAA, BACK cull",
                W=cfg["W"], H=cfg["H"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--extra-configs", default=None, help="comma list of further configs in the same line (short runs); default c2,c4,c5 at N=1, c5 at N>1")
    ap.add_argument("--no-exact-sampler", "--no-fast-sampler", dest="no_exact_sampler", action="store_true",
                    help="skip the extra run with the library's default (exact) sampler")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emulate-rank", default=None, help="analysis aid, 1 GPU: 'R/N' renders only the tiles rank R of N owns (per-rank kernel times, no exchange)")
    ap.add_argument("--no-distributed-setup", action="store_true", help="A/B (N > 1): every rank assembles every triangle")
    ap.add_argument("--force-distributed-setup", action="store_true", help="A/B (N > 1): split the triangle setup whatever the rank / triangle count")
    ap.add_argument("--no-distributed-readback", action="store_true", help="A/B (N > 1): e2e reads the frame back through rank 0 only")
    ap.add_argument("--ring", type=int, default=2, choices=[1, 2], help="A/B (N > 1, peer mode): frames in flight in the owner's framebuffer (2 = the peers render frame k + 1 while the owner releases frame k)")
    ap.add_argument("--no-preclear", action="store_true", help="A/B (N > 1, peer mode): plain frame release, every rank sends all of its tiles")
    ap.add_argument("--tile-chunk", type=int, default=0, choices=[0, 32, 64], help="A/B: big-queue entries per staged chunk of the tile kernel (0 = automatic)")
    ap.add_argument("--no-grid-hint", action="store_true", help="A/B: tile kernel grid of one CTA per tile instead of the measured CTA count")
    ap.add_argument("--no-tma", action="store_true", help="A/B: framebuffer stores / clears by per-thread STG instead of TMA tensor copies")
    ap.add_argument("--mgpu", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = tiles stored straight into rank 0's framebuffer over NVLink; nccl = pack + gather + unpack")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    cfg_name = args.config or "c3"   # the configuration BASELINE.json quotes the roofline target on (4K, 8x MSAA, textured)

    if args.impl == "reference":
        if rank != 0:
            return
        res = run_reference(args, cfg_name, steps=min(args.steps, 20), warmup=min(max(args.warmup, 1), 3))
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libifrit_ref.so missing (build needs /root/reference)"}))
            return
        cfg = CONFIGS[cfg_name]
        print(json.dumps({
            "impl": "reference", "metric": "Mtris/s", "value": res["value"], "unit": "Mtris/s", "n_gpus": args.gpus,
            "steps": min(args.steps, 20), "warmup": min(max(args.warmup, 1), 3), "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (AVX2, rcpps)",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": cfg_name, "triangles": res["ntri"], "resolution": [res["W"], res["H"]]},
            "fps": 1e3 / res["ms_per_step"],
            "cpu_baseline": {"value": res["value"], "unit": "Mtris/s", "cores": res["cores"], "threads": res["threads"],
                             "kind": "reference", "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = run_ours(args, cfg_name)
    # Further configurations in the same line (short runs): N = 1: the other single-GPU BASELINE configs and C5;
    # N > 1: C5, the configuration BASELINE.json quotes the multi-GPU partition on — every rank takes part.
    default_extras = "c2,c4,c5" if world == 1 else "c5"
    extra_names = [c for c in (default_extras if args.extra_configs is None else args.extra_configs).split(",") if c and c != cfg_name]
    extras = {}
    keys = ("value", "ms_per_step", "frame_latency_ms", "fps", "gfrags_per_s", "kernel_ms", "kernel_ms_unpipelined", "roofline",
            "frame_roofline", "e2e", "gpu_launches", "n_gpus", "frame_hash_equals_1gpu")
    for name in extra_names:
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup = max(5, min(args.steps, 20)), 3
        try:
            o = run_ours(sub, name, do_e2e=(world == 1 and name in ("c2", "c3")))
            if rank == 0:
                extras[name] = {k: o[k] for k in keys if k in o}
                extras[name]["workload"] = o["config"]["workload"]
        except Exception as ex:   # an extra config must never take the headline down
            extras[name] = {"error": repr(ex)}
    if world == 1 and CONFIGS[cfg_name]["fs"] == "textured" and not args.no_exact_sampler:
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup = max(5, min(args.steps, 100)), 3
        try:
            o = run_ours(sub, cfg_name, do_e2e=False, fast_sampler=False)
            out["exact_sampler"] = {"ms_per_step": o["ms_per_step"], "value": o["value"], "kernel_ms": o["kernel_ms"],
                                    "roofline_frac": o["roofline"]["frac"],
                                    "note": "same workload with the library's default sampler (software fp32 lerps, the oracle's arithmetic op for "
                                            "op); the headline opts in to the FAST sampler (iftrcuSetSamplerHardwareFilter(1), include/ifrit_b200.h), "
                                            "both are parity-tested at this size (tests/test_gpu_parity.py::test_full_size_c2_c3_vs_oracle)"}
        except Exception as ex:
            out["exact_sampler"] = {"error": repr(ex)}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    if extras:
        out["extra_configs"] = extras
    if world == 1:
        if not args.no_cpu_baseline:
            try:
                res = run_reference(args, cfg_name, steps=10, warmup=2)
                if res:
                    out["cpu_baseline"] = {"value": res["value"], "unit": "Mtris/s", "cores": res["cores"],
                                           "threads": res["threads"], "kind": "reference", "ms_per_frame": res["ms_per_step"],
                                           "sample": res["sample"]}
            except Exception as ex:
                out["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
