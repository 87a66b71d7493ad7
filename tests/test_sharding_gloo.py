"""N>1 host logic on CPU: world_size-2 gloo processes shard (a) a batch of frames and (b) the cells of a grid canvas,
convert their units with the oracle (the GPU box uses the CUDA library through the same code), and the root checks that
the reassembled result is byte-identical to converting everything in one process (cell-by-cell == canvas)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from libavif_b200 import abi, sharding
    from libavif_b200.images import RGBImage, YUVImage
    from oracle import loader
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = loader.load_oracle()
    A = abi
    try:
        # ---- (a) frame batch: frame f -> rank f mod G, no communication except the final hash gather ----
        n_frames, w, h = 5, 96, 64
        my_hashes = {}
        for f in sharding.units_of(rank, n_frames, world):
            rs = np.random.RandomState(100 + f)
            img = YUVImage(w, h, 10, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_LIMITED, A.MC_BT709)
            for p in img.present_planes():
                pw, ph = img.plane_dims(p)
                img.set_plane_samples(p, rs.randint(0, 1024, size=(ph, pw)))
            rgb = RGBImage(w, h, 16, A.AVIF_RGB_FORMAT_RGB)
            assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
            my_hashes[f] = int(lib.raw.oracleFnv1a64(0xcbf29ce484222325, rgb.buf.ptr, rgb.buf.nbytes))
        gathered = [None] * world
        dist.all_gather_object(gathered, my_hashes)

        # ---- (b) grid: 3x2 cells of a 4:2:0 canvas, RGB -> YUV, scatter / convert / gather ----
        cw_, ch_, cols, rows = 200, 150, 3, 2
        rects = sharding.split_grid(cw_, ch_, cols, rows, A.AVIF_PIXEL_FORMAT_YUV420)
        canvas = None
        if rank == 0:
            rs = np.random.RandomState(7)
            canvas = RGBImage(cw_, ch_, 8, A.AVIF_RGB_FORMAT_RGBA)
            canvas.set_payload(rs.randint(0, 256, size=(ch_, cw_ * 4)))

        def cell_pixels(u):
            r = rects[u]
            return torch.from_numpy(canvas.buf.host()[r.y:r.y + r.height, r.x * 4:(r.x + r.width) * 4].copy())

        def send_cell(u, dst):
            t = cell_pixels(u)
            if dst != 0:
                dist.send(t, dst)
            return t

        def recv_cell(u, src):
            r = rects[u]
            t = torch.empty((r.height, r.width * 4), dtype=torch.uint8)
            dist.recv(t, src)
            return t

        def convert_cell(u, t):
            r = rects[u]
            rgb = RGBImage(r.width, r.height, 8, A.AVIF_RGB_FORMAT_RGBA)
            rgb.set_payload(t.numpy())
            img = YUVImage(r.width, r.height, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True)
            assert lib.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
            return [torch.from_numpy(img.plane_samples(p).copy()) for p in range(4)]

        def send_result(u, planes, dst):
            for t in planes:
                dist.send(t.contiguous(), dst)

        def recv_result(u, src):
            r = rects[u]
            shapes = [(r.height, r.width), ((r.height + 1) // 2, (r.width + 1) // 2), ((r.height + 1) // 2, (r.width + 1) // 2), (r.height, r.width)]
            out = []
            for s in shapes:
                t = torch.empty(s, dtype=torch.uint8)
                dist.recv(t, src)
                out.append(t)
            return out

        res = sharding.scatter_convert_gather(dist, rank, world, rects, send_cell, recv_cell, convert_cell, send_result, recv_result)
        ok = True
        if rank == 0:
            # frames: every frame converted exactly once, by the expected owner
            merged = {}
            for r_, d in enumerate(gathered):
                for f, hsh in d.items():
                    assert sharding.owner_of(f, world) == r_
                    merged[f] = hsh
            assert sorted(merged) == list(range(n_frames))
            # grid: tiled planes == whole-canvas conversion
            whole = YUVImage(cw_, ch_, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True)
            assert lib.avifImageRGBToYUV(whole.ptr(), canvas.ptr()) == 0
            for u, r in enumerate(rects):
                for p in range(4):
                    sx = 1 if p in (1, 2) else 0
                    ref = whole.plane_samples(p)[r.y >> sx:(r.y >> sx) + ((r.height + sx) >> sx), r.x >> sx:(r.x >> sx) + ((r.width + sx) >> sx)]
                    ok = ok and bool((res[u][p].numpy() == ref).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_frame_and_grid_sharding_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results)


def test_split_rules_match_reference_semantics():
    from libavif_b200 import abi, sharding
    # 4x4 grid of a 7680x4320 4:2:0 canvas (config C4): 1920x1080 cells, all even
    rects = sharding.split_grid(7680, 4320, 4, 4, abi.AVIF_PIXEL_FORMAT_YUV420)
    assert len(rects) == 16 and all(r.width == 1920 and r.height == 1080 for r in rects)
    # odd split on a subsampled image is bumped to even; the last cell is clipped
    rects = sharding.split_grid(1001, 333, 3, 2, abi.AVIF_PIXEL_FORMAT_YUV420)
    assert [r.width for r in rects[:3]] == [334, 334, 333] and rects[0].height == 168 and rects[3].height == 165
    assert all(r.x % 2 == 0 and r.y % 2 == 0 for r in rects)
    # 4:4:4 keeps odd cells
    assert sharding.split_grid(1001, 333, 3, 2, abi.AVIF_PIXEL_FORMAT_YUV444)[0].width == 334
    assert sharding.split_grid(999, 333, 3, 1, abi.AVIF_PIXEL_FORMAT_YUV444)[0].width == 333
    with pytest.raises(ValueError):
        sharding.split_grid(100, 100, 3, 1, abi.AVIF_PIXEL_FORMAT_YUV420)   # 64-px minimum cells: the third would be off-canvas
    assert sharding.units_of(1, 7, 4) == [1, 5] and sharding.owner_of(6, 4) == 2
