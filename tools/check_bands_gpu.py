"""Run under torchrun on N GPUs: the band-partitioned frame (render, blur-halo exchange, local blur, gather)
must equal the single-GPU frame bit for bit. Prints OK/FAIL per scene on rank 0, exits non-zero on mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_bands_gpu.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from raytracingdiffusioncurves_b200 import api, distributed as rd  # noqa: E402

XML = os.path.join(ROOT, "tests", "golden", "xmls")
CASES = [("arch.xml", 640, 360, 32), ("DiffusionCurvePack/lady_bug.xml", 512, 384, 16), ("DiffusionCurvePack/face.xml", 256, 64, 8),
         ("PortalDemo.xml", 320, 242, 16)]


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

        def make(b=0, e=h):
            return api.default_frame_params(w, h, rpp, zoom_factor=zoom, row_begin=b, row_end=e, frame=3)

        halo = rd.halo_rows(host.max_blur(2))
        plan = rd.StripPlan(h, w, world, rank, halo)
        bands = rd.FrameBuffers(plan, dev)
        render_strips, blur_rows = api.cuda_callbacks(scene, make, stream)
        frame = rd.render_frame(bands, render_strips, blur_rows, use_blur=True)
        torch.cuda.synchronize()
        # the same frame through peer memory (render kernel stores into the consumers' frames)
        peer_frame, peer_note = None, "peer path unavailable"
        try:
            peers = rd.PeerFrameBuffers(plan, dev)
            render_to, blur_rows_p = api.cuda_peer_callbacks(scene, make, stream)
            for _ in range(2):  # twice: buffers are reused from frame to frame
                peer_frame = rd.render_frame_peer(peers, render_to, blur_rows_p, use_blur=True)
            torch.cuda.synchronize()
            peer_note = "peer path"
        except Exception as exc:  # noqa: BLE001
            peer_note = f"peer path unavailable: {type(exc).__name__}: {exc}"
        if rank == 0:
            image = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
            sigma = torch.empty((h, w), dtype=torch.float32, device=dev)
            scratch = torch.empty_like(image)
            want = torch.empty_like(image)
            scene.render(make(0, h), image.data_ptr(), sigma.data_ptr(), stream)
            if halo > 0:
                api.gaussian_blur(want.data_ptr(), image.data_ptr(), sigma.data_ptr(), scratch.data_ptr(), w, h, 0, h, 0, stream)
            else:
                want = image
            torch.cuda.synchronize()
            same = torch.equal(frame[:h, :, :3].contiguous().view(torch.int32), want[..., :3].contiguous().view(torch.int32))
            print(f"{'OK  ' if same else 'FAIL'} {name} {w}x{h}@{rpp} world={world} halo={halo}", flush=True)
            failures += 0 if same else 1
            if peer_frame is not None:
                same = torch.equal(peer_frame[..., :3].contiguous().view(torch.int32), want[..., :3].contiguous().view(torch.int32))
                print(f"{'OK  ' if same else 'FAIL'} {name} {peer_note}", flush=True)
                failures += 0 if same else 1
            else:
                print(f"SKIP {name} {peer_note}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
