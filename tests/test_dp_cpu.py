"""world_size-2 gloo test (CPU): the data-parallel decomposition the engine uses -- rank-partitioned batch and
noise, gradient mean over ranks before every optimiser update, global-batch PID / MMD scalars -- reproduces
the single-process step on the concatenated batch."""
import socket

import pytest
import torch.multiprocessing as mp

from tests import dp_worker


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("algo", ["bc", "bcql", "bearl", "cpq", "cdt"])
def test_two_rank_equivalence_gloo(algo):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=dp_worker._mp_entry, args=(r, 2, algo, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert ok
