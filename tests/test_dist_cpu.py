"""Multi-rank logic on CPU: world_size 2, gloo.  The per-rank block computation (CUDA kernels in production) is
replaced by the oracle here; what is tested is the sharding, the K/V all-gather and the stitching: the sharded
matrix must equal the single-process matrix bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, ret):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from diffsim_b200 import scoring, synth
    from oracle import aas_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = synth.SynthModel(1, 2, 32, 40, seed=2334)
        images, _ = synth.make_styles(m, n_images, 1, torch.float32, seed=8)

        def oracle_block(local, k_all, v_all, similarity, scale):
            out = torch.zeros(local.n_images, k_all.shape[0], dtype=torch.float64)
            for i in range(local.n_images):
                for j in range(k_all.shape[0]):
                    out[i, j] = O.aas_directional(local.q[i], local.k[i], local.v[i], k_all[j], v_all[j], similarity)
            return out

        scoring._matrix_block = oracle_block
        r0, r1 = scoring.row_block(n_images, rank, world)
        local = scoring.QKVCache.from_images(images[r0:r1]) if r1 > r0 else None
        if local is None:  # a rank may own no rows
            B, H, S, D = 1, 2, 32, 40
            local = scoring.QKVCache.empty(0, B, H, S, D, torch.float32, "cpu")
        full = scoring.aas_matrix_sharded(local, "cosine")
        if rank == 0:
            ref = O.aas_matrix([i[0] for i in images], [i[1] for i in images], [i[2] for i in images])
            ret["equal"] = bool(torch.equal(full, ref))
            ret["shape"] = tuple(full.shape)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_images", [(2, 5), (2, 4), (3, 2)])
def test_sharded_matrix_equals_single_process(world, n_images):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_images, ret), nprocs=world, join=True)
    assert ret["shape"] == (n_images, n_images)
    assert ret["equal"]
