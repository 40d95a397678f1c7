"""N > 1 host logic on CPU: world_size-2 gloo. Each rank renders its shard of the sample range (the oracle stands in
for the device renderer — tests may use it), the accumulators are combined with the path's single collective,
reduce(SUM) to rank 0, and the result equals the single-rank render of the full range."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp


def _worker(rank, world, port, out_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle as O
    import ptb200

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scene = ptb200.load_file(os.path.join(root, "scenes", "rtweekend1.ssml"))
    o = O.OracleScene(scene)
    total_spp, w, h = 6, 48, 27
    off, n = ptb200.shard_samples(total_spp, rank, world)
    acc, _, _ = o.render(w, h, n, 1, seed=4, sample_offset=off, threads=1)
    t = torch.from_numpy(acc.reshape(-1).copy())
    img = ptb200.reduce_accumulators(t, total_spp, dst=0)
    if rank == 0:
        np.save(out_path, img.numpy())
    else:
        assert img is None
    dist.destroy_process_group()


def test_spp_sharding_reduce_equals_single_rank(ptb, orc, rtweekend1, tmp_path):
    out = str(tmp_path / "img.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    full, _, _ = orc.OracleScene(rtweekend1).render(48, 27, 6, 1, seed=4, threads=1)
    assert np.allclose(got, full.reshape(-1) / 6, atol=1e-6)


def test_shard_samples_tiles_the_range(ptb):
    for total in (0, 1, 7, 256, 4096):
        for world in (1, 2, 3, 8):
            parts = [ptb.shard_samples(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == total
            for (a, n), (b, _) in zip(parts, parts[1:]):
                assert a + n == b
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1


def test_c_abi_shard_helper_matches(ptb):
    """ptb_shard_samples (what ptb_render_multi uses inside the library) == the Python helper the torchrun path uses."""
    import ctypes as C
    for total in (0, 1, 7, 256, 4096):
        for world in (1, 2, 3, 8):
            for r in range(world):
                a, b = C.c_uint32(), C.c_uint32()
                ptb._lib.lib.ptb_shard_samples(total, 11, r, world, C.byref(a), C.byref(b))
                assert (a.value - 11, b.value) == ptb.shard_samples(total, r, world)
