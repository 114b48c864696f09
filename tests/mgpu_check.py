"""Run under torchrun (one rank per GPU): the sort-first frame assembled on rank 0 — peer-memory stores over
NVLink or NCCL gather — must be bit-identical to the single-GPU frame.  Prints MGPU_OK."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ifrit_v2_b200  # noqa: E402,F401
from ifrit_v2_b200 import renderer as R, scenes, multigpu as MG  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H = 1000, 600
    vb, idx = scenes.mesh_sphere(n=48)
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    tex = scenes.texture_checker(size=128, seed=7)
    ok = True
    modes = os.environ.get("MGPU_MODES", "peer,peer+preclear,peer+preclear+dsetup,peer+ring,peer+preclear+ring,nccl").split(",")
    for mode in modes:
        for msaa in (1, 8):
            r = R.TileRasterRendererCuda(local)
            r.Init()
            stream = torch.cuda.Stream(device=dev)
            torch.cuda.set_stream(stream)
            r.setStream(stream.cuda_stream)
            r.width, r.height = W, H
            if mode == "nccl":
                r.bindFrameBuffer(None, None, W, H)
            sf = MG.SortFirstRenderer(r, dist, dev, mode=mode.split("+")[0], clears_every_frame="preclear" in mode, ring=2 if "ring" in mode else 1)
            R.setup_scene(r, vb, idx, R.VS_MVP, R.FS_TEXTURED, 2, u, texture=tex, mip_levels=7,
                          sampler=dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0))
            r.setMsaaSamples(msaa)
            r.SetCullMode(R.IF_CULL_MODE_BACK)
            r.setClearValues([(0.1, 0.2, 0.3, 1.0)], 255.0)
            nframes = 5 if "ring" in mode else 3       # ring: both buffers used twice, pre-cleared releases of both in effect
            for f in range(nframes):             # frames 2 and 3 of "peer+preclear" run with the empty tiles kept off NVLink
                r.clear()
                if "ring" in mode:      # earlier frames differ from the last one (other faces culled): a stale or wrong buffer would show
                    r.SetCullMode(R.IF_CULL_MODE_BACK if f == nframes - 1 else R.IF_CULL_MODE_FRONT)
                if mode.endswith("dsetup"):          # triangle setup split over the ranks, owner masks all-gathered
                    sf.distribute_setup(len(idx) // 3)
                r.drawElements()
                sf.gather(hold=(f == nframes - 1))   # the last frame is held until it has been read (a release lets it be overwritten)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                got = r.readColor()
                sf.release()
                ref = R.TileRasterRendererCuda(local)
                ref.Init()
                ref.bindFrameBuffer(None, None, W, H)
                R.setup_scene(ref, vb, idx, R.VS_MVP, R.FS_TEXTURED, 2, u, texture=tex, mip_levels=7,
                              sampler=dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0))
                ref.setMsaaSamples(msaa)
                ref.SetCullMode(R.IF_CULL_MODE_BACK)
                ref.setClearValues([(0.1, 0.2, 0.3, 1.0)], 255.0)
                ref.clear()
                ref.drawElements()
                ref.synchronize()
                want = ref.readColor()
                same = np.array_equal(got.view(np.uint32), want.view(np.uint32))
                print(f"mode={mode} msaa={msaa} world={world} identical={same} covered={(want[..., 3] == 1.0).mean():.3f}", flush=True)
                ok = ok and same
                ref.close()
            dist.barrier()
            r.close()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
