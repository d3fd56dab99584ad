"""Worker for tests/test_sharded.py: one rank of a prime-sharded key switch (gloo + simulator on CPU,
nccl + CUDA on GPUs).  Prints 'RANK r OK' when the rows this rank owns match the oracle bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    backend, cfg_s = sys.argv[1], sys.argv[2]
    cfg = tuple(int(x) for x in cfg_s.split(","))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import orc
    import pyoracle as po
    from helib_b200.engine import Engine, load_library
    from helib_b200.sharded import ShardedKeySwitch
    if backend == "sim":
        from conftest import build_sim
        lib = load_library(build_sim())
        dist.init_process_group("gloo")
        device = "cpu"
    elif backend == "cuda1":
        # world_size ranks sharing ONE GPU (the driver's 1-GPU box): real library, gloo for the plumbing
        # (NCCL refuses two ranks on one device), CUDA-IPC peer mappings for the p2p stores.
        lib = load_library()
        local = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
        device = "cuda:0"
    else:
        lib = load_library()
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        device = f"cuda:{local}"
    ch = po.build_mod_chain(*cfg)
    psis = [po.find_psi(q, ch.m) for q in ch.primes]
    O = orc.Oracle(ch.phim, ch.m, ch.primes, psis, ch.digits, ch.special, nthreads=4)
    E = Engine(ch.m, ch.primes, psis, ch.digits, ch.special, device=0 if backend in ("sim", "cuda1") else int(os.environ.get("LOCAL_RANK", rank)), lib=lib)
    if backend != "sim":
        E.set_stream(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(123)           # same data on every rank
    full = ch.ctxt + ch.special
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    p = 1 if ch.p == -1 else ch.p ** ch.r
    p2p = len(sys.argv) > 3 and sys.argv[3] == "p2p"
    gsz = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # > 0: independent prime-sharded groups of gsz ranks side by side
    group = None
    if gsz and gsz < world:
        groups = [dist.new_group(list(range(i, i + gsz))) for i in range(0, world, gsz)]
        group = groups[rank // gsz]
    KS = ShardedKeySwitch(E, ch.ctxt, ch.special, ch.digits, device=device, p2p=p2p, group=group)
    own_full = KS.owned(full)
    EA = [E.poly(evk_a[i], own_full) for i in range(nd)]   # evk sharded identically: only owned rows uploaded
    EB = [E.poly(evk_b[i], own_full) for i in range(nd)]
    ok = True
    sets = [ch.ctxt, ch.ctxt[:-1]]
    if backend == "sim" and len(ch.digits) >= 3:
        sets.append([i for i in ch.ctxt if i not in ch.digits[1]])    # a hole: digit 1 has no live prime
    for S in sets:
        nit = 2
        cs = [[O.random(rng, S) for _ in range(3)] for _ in range(nit)]
        oS = KS.owned(S)
        C0, C1, C2 = ([E.poly(c[k], oS) for c in cs] for k in range(3))
        Sp = KS.relinearize(C0, C1, C2, S, EA, EB)
        KS.mod_down(C0 + C1, Sp, S, p)
        if backend != "sim":
            torch.cuda.synchronize()
        for it in range(nit):
            r0, r1 = O.relinearize(cs[it][0], cs[it][1], cs[it][2], S, evk_a, evk_b)
            O.scale_down(r0, Sp, S, p); O.scale_down(r1, Sp, S, p)
            g0, g1 = C0[it].download(oS), C1[it].download(oS)
            ok &= bool((g0[oS] == r0[oS]).all() and (g1[oS] == r1[oS]).all())
    dist.barrier()
    print(f"RANK {rank} {'OK' if ok else 'MISMATCH'} owned={len(own_full)}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
