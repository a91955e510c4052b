"""Two-GPU checks (NCCL): the row-block sharded all-pairs matrix equals the single-GPU matrix bit for bit, and the
pair-sharded scoring equals the unsharded one.  Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, ret):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    from diffsim_b200 import ops, scoring, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        m = synth.SynthModel(2, 4, 256, 64, seed=2334)
        images, labels = synth.make_styles(m, (n_images + 1) // 2, 2, torch.float16, seed=8)
        images = images[:n_images]
        r0, r1 = scoring.row_block(n_images, rank, world)
        local = scoring.QKVCache.from_images(images[r0:r1], dev)
        ev = {}
        full = scoring.aas_matrix_sharded(local, "cosine", timings=ev)
        torch.cuda.synchronize()
        # pair list sharded over the ranks, scores gathered (no data-path collective)
        pairs = [(i, (i * 7 + 3) % n_images) for i in range(n_images)]
        p0, p1 = scoring.row_block(len(pairs), rank, world)
        whole = scoring.QKVCache.from_images(images, dev)
        mine = scoring.score_pairs(whole, pairs[p0:p1])
        parts = [torch.empty(scoring.row_block(len(pairs), r, world)[1] - scoring.row_block(len(pairs), r, world)[0],
                             dtype=torch.float32, device=dev) for r in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            ref = ops.aas_matrix(whole.q, whole.k, whole.v, whole.k, whole.v, "cosine")
            ret["matrix_equal"] = bool(torch.equal(full, ref))
            ret["pairs_equal"] = bool(torch.equal(torch.cat(parts), scoring.score_pairs(whole, pairs)))
            ret["shape"] = tuple(full.shape)
            ret["events"] = sorted(ev.keys())
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 7])
def test_sharded_matrix_and_pairs_on_two_gpus(n_images):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp

    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), n_images, ret), nprocs=2, join=True)
    assert ret["shape"] == (n_images, n_images)
    assert ret["matrix_equal"] and ret["pairs_equal"]
    assert ret["events"] == ["block_done", "end", "exchange_done", "own_done", "start"]
