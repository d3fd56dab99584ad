"""Prime-sharded key switching: world_size-2/3 over gloo with the CPU kernel simulator,
and over NCCL on real GPUs (-m gpu, needs >= 2 devices)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mp", "sharded_worker.py")


def run(backend, world, cfg, port, mode="gather", group_size=0):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, backend, cfg, mode, str(group_size)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    for k in range(world):
        assert f"RANK {k} OK" in out, out[-3000:]


@pytest.mark.parametrize("world,cfg,port", [(2, "64,257,1,120,2", 29611), (3, "4096,17,1,160,3", 29612)])
def test_sharded_keyswitch_gloo_sim(sim_lib, world, cfg, port):
    run("sim", world, cfg, port)


def test_sharded_keyswitch_in_independent_groups_gloo_sim(sim_lib):
    """Four ranks as two prime-sharded groups of two (the grouping bench.py reports next to the pure sharding)."""
    run("sim", 4, "64,257,1,120,2", 29618, "gather", group_size=2)


def test_owner_map_balances_digits():
    import pyoracle as po
    from helib_b200.sharded import owner_map
    ch = po.build_mod_chain(1 << 17, -1, 1, 1700, 2)
    for world in (2, 4, 8):
        own = owner_map(ch.ctxt, ch.special, world)
        for cls in (ch.ctxt, ch.special) + tuple(ch.digits):
            counts = [sum(1 for i in cls if own[i] == r) for r in range(world)]
            assert max(counts) - min(counts) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("mode,port", [("gather", 29616), ("p2p", 29617)])
def test_sharded_keyswitch_two_ranks_one_gpu(cuda_lib, mode, port):
    """world_size 2 on ONE device with the real library (gloo plumbing; p2p = CUDA-IPC peer stores from the producing
    kernel): hb_conv_make_y / hb_conv_make_y_bcast / hb_conv_from_y parity without needing a second GPU."""
    run("cuda1", 2, "8192,257,1,160,2", port, mode)
    if mode == "p2p":
        run("cuda1", 2, "131072,257,1,230,2", port + 2, mode)


def _backend():
    """NCCL over two devices when the box has them, otherwise the same two ranks on one device (gloo plumbing)."""
    import torch
    return "cuda" if torch.cuda.device_count() >= 2 else "cuda1"


@pytest.mark.gpu
def test_sharded_keyswitch_nccl(cuda_lib):
    run(_backend(), 2, "8192,257,1,160,2", 29613)


@pytest.mark.gpu
def test_sharded_keyswitch_p2p_stores(cuda_lib):
    """Same circuit with the fused make-y + peer-store kernel (CUDA IPC; NVLink between two GPUs) instead of all_gather."""
    run(_backend(), 2, "8192,257,1,160,2", 29614, "p2p")
    run(_backend(), 2, "131072,257,1,230,2", 29615, "p2p")
