"""Run under torchrun on N GPUs (one process per GPU): the frame split over the ranks through the C ABI's peer frames
— CUDA IPC handles, render kernel stores into the consumers' frames, barrier kernels, per-rank copies into one shared
pinned host frame — must equal the single-GPU frame bit for bit, for the device consumer and the host consumer, and so
must the NCCL form of the same plan (distributed.render_frame). Prints OK/FAIL per scene on rank 0, exits non-zero on
mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_bands_gpu.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from raytracingdiffusioncurves_b200 import api, distributed as rd  # noqa: E402

XML = os.path.join(ROOT, "tests", "golden", "xmls")
CASES = [("arch.xml", 640, 360, 32), ("DiffusionCurvePack/lady_bug.xml", 512, 384, 16), ("DiffusionCurvePack/face.xml", 256, 64, 8),
         ("PortalDemo.xml", 320, 242, 16), ("arch.xml", 1920, 1080, 128), ("DiffusionCurvePack/lady_bug.xml", 960, 540, 128)]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream
    failures = 0
    for name, w, h, rpp in CASES:
        host = api.HostScene.from_xml_file(os.path.join(XML, name))
        scene = api.Scene(host.arrays, None, stream)
        zoom = host.arrays.image_height / h

        def make(frame=3, b=0, e=h):
            # units per tile pinned: the whole frame and a rank's share would otherwise pick different summation orders
            # (8 = group mode where a launch has 128 rays per pixel; fewer rays clamp it, the same way on both sides)
            return api.default_frame_params(w, h, rpp, zoom_factor=zoom, row_begin=b, row_end=e, frame=frame, units_per_tile=8)

        halo = host.halo_rows(2)
        scene.reserve(make(), False, stream)
        # ---- NCCL form ----
        plan = rd.StripPlan(h, w, world, rank, halo)
        bands = rd.FrameBuffers(plan, dev)
        render_strips, blur_rows = api.cuda_callbacks(scene, make, stream)
        frame = rd.render_frame(bands, render_strips, blur_rows, use_blur=True)
        torch.cuda.synchronize()
        # ---- C ABI peer frames ----
        peers = api.PeerFrames(w, h, rank, world)
        mine = torch.frombuffer(bytearray(peers.export_handles()), dtype=torch.uint8).to(dev)
        every = torch.empty((world * api.PEER_HANDLE_BYTES,), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(every, mine)
        peers.connect_ipc(every.cpu().numpy().tobytes())
        dist.barrier()
        ptr = 0
        for f in range(3):  # three frames: both frame buffers are used, one of them twice
            ptr = peers.render_frame(scene, make(f + 1), True, halo, stream)
        torch.cuda.synchronize()
        peers.status()
        names = [f"/rdc_check_{os.getpid()}_{k}" for k in range(2)] if rank == 0 else [None, None]
        dist.broadcast_object_list(names, src=0)
        hosts = [api.HostFrame(nm, h * w * 16, True) for nm in names] if rank == 0 else []
        dist.barrier()
        if rank != 0:
            hosts = [api.HostFrame(nm, h * w * 16, False) for nm in names]
        for f in range(4):
            peers.frame_to_host(scene, make(f), True, halo, hosts[f % 2].ptr, stream)
        peers.wait()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            image = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
            sigma = torch.empty((h, w), dtype=torch.float32, device=dev)
            scratch = torch.empty_like(image)

            def single(frame_no):
                want = torch.empty_like(image)
                scene.render(make(frame_no), image.data_ptr(), sigma.data_ptr(), stream)
                if halo > 0:
                    api.gaussian_blur(want.data_ptr(), image.data_ptr(), sigma.data_ptr(), scratch.data_ptr(), w, h, 0, h, 0, stream)
                else:
                    want.copy_(image)
                torch.cuda.synchronize()
                return want[..., :3].contiguous().view(torch.int32)

            def report(ok, what):
                nonlocal failures
                print(f"{'OK  ' if ok else 'FAIL'} {name} {w}x{h}@{rpp} world={world} halo={halo}: {what}", flush=True)
                failures += 0 if ok else 1

            report(torch.equal(frame[:h, :, :3].contiguous().view(torch.int32), single(3)), "NCCL form")
            got = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
            ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(got.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(h * w * 16), 3)
            report(torch.equal(got[..., :3].contiguous().view(torch.int32), single(3)), "peer frames, device consumer (third frame)")
            for f in (2, 3):
                host_frame = torch.from_numpy(hosts[f % 2].numpy((h, w, 4)).copy()).to(dev)
                report(torch.equal(host_frame[..., :3].contiguous().view(torch.int32), single(f)), f"peer frames, host consumer (frame {f})")
        dist.barrier()
        for hf in hosts:
            hf.close()
        peers.close()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
