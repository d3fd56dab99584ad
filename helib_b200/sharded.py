"""Prime-sharded key switching across GPUs (BASELINE config 4).

One process per GPU.  Rows are sharded by RNS prime index (round-robin inside the ctxt primes and
inside the special primes, so every rank owns an equal share of every digit); the evaluation-key rows
are sharded identically and never move.  Residues cross shards at exactly two points, each one
all-gather over torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests):

  1. breakIntoDigits: the y_j = iNTT(row_j)*(Q_D/q_j)^-1 rows of each digit (l rows in total),
  2. mod-down: the y rows of the K special primes of both parts.

Everything else (transforms, exact CRT to the owned target rows, mixed-radix updates, evk inner
product) is local.  torch is plumbing here: tensors own the exchange buffers, torch.distributed
carries them; all arithmetic is in the engine's kernels (hb_conv_make_y / hb_conv_from_y).

Reference semantics: Ctxt::reLinearize / keySwitchPart / keySwitchDigits (src/Ctxt.cpp:191-230,720-842),
DoubleCRT::breakIntoDigits (src/DoubleCRT.cpp:479-561), Ctxt::modDownToSet (src/Ctxt.cpp:393-562).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def owner_map(ctxt, special, world):
    """prime index -> owning rank: round-robin within each class."""
    own = {}
    for cls in (ctxt, special):
        for k, i in enumerate(cls):
            own[i] = k % world
    return own


class ShardedKeySwitch:
    def __init__(self, eng, ctxt, special, digits, rank=None, world=None, device="cuda", p2p=False, group=None):
        """group (optional): a torch.distributed process group -- the rows are sharded over ITS ranks only (rank / world are then
        the rank and size inside the group), so a box can run several independent prime-sharded groups side by side."""
        self.E = eng
        self.group = group
        if group is not None:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.ctxt, self.special, self.digits = list(ctxt), list(special), [list(d) for d in digits]
        self.owner = owner_map(self.ctxt, self.special, self.world)
        self.device = device
        self._bufs = {}
        # p2p: the y rows are stored straight into the peers' buffers by the producing kernel (CUDA IPC mappings
        # over NVLink) instead of pack -> all_gather -> unpack; needs real GPUs, one process per GPU.
        self.p2p = bool(p2p) and self.world > 1
        self._flag = torch.zeros(1, dtype=torch.int32, device=device) if self.p2p else None

    def owned(self, idx):
        return [i for i in idx if self.owner[i] == self.rank]

    # ---- exchange buffers: torch owns the memory, the engine aliases it
    def _ybuf(self, key):
        if key not in self._bufs:
            if self.p2p:   # engine-owned buffer, exported to / imported from every peer (collective: same key order on all ranks)
                mine = self.E.poly()
                handles = [None] * self.world
                dist.all_gather_object(handles, self.E.ipc_export(mine), group=self.group)
                peers = [self.E.ipc_open(h) for r, h in enumerate(handles) if r != self.rank]
                self._bufs[key] = (None, mine, peers)
            else:
                t = torch.zeros((self.E.np, self.E.N), dtype=torch.int64, device=self.device)
                self._bufs[key] = (t, self.E.wrap(t.data_ptr()), None)
        return self._bufs[key]

    def _exchange(self, polys, D, ys):
        """y rows of the source set D: produced for the owned rows, present on every rank afterwards."""
        E = self.E
        if self.p2p:
            npeer = self.world - 1
            peers = [[y[2][p] for y in ys] for p in range(npeer)]
            E.conv_make_y_bcast(polys, D, self.owned(D), [y[1] for y in ys], peers)
            dist.all_reduce(self._flag, group=self.group)          # stream-ordered cross-rank barrier: all peers' stores have landed
        else:
            E.conv_make_y(polys, D, self.owned(D), [y[1] for y in ys])
            self._all_gather_rows([y[0] for y in ys], D)

    def _all_gather_rows(self, tensors, D):
        """After this call every rank holds rows D of every tensor in `tensors` (each rank contributed the
        rows it owns).  One all_gather of a packed [n_items, max_owned, N] buffer."""
        if self.world == 1:
            return
        key = ("rows", tuple(D))
        if key not in self._bufs:   # index tensors are built once (host->device copies are not graph-capturable)
            per = [[i for i in D if self.owner[i] == r] for r in range(self.world)]
            self._bufs[key] = (per, [torch.tensor(p, dtype=torch.int64, device=self.device) if p else None for p in per])
        per, rowt = self._bufs[key]
        mx = max(len(p) for p in per)
        mine = per[self.rank]
        skey = ("send", len(tensors), mx)
        if skey not in self._bufs:
            self._bufs[skey] = (torch.zeros((len(tensors), mx, self.E.N), dtype=torch.int64, device=self.device),
                                torch.empty(self.world * len(tensors) * mx * self.E.N, dtype=torch.int64, device=self.device))
        send, flat = self._bufs[skey]
        if mine:
            for k, t in enumerate(tensors):
                send[k, :len(mine)] = t.index_select(0, rowt[self.rank])
        dist.all_gather_into_tensor(flat, send.view(-1), group=self.group)
        recv = flat.view((self.world,) + tuple(send.shape))
        for r in range(self.world):
            if r == self.rank or not per[r]:
                continue
            for k, t in enumerate(tensors):
                t.index_copy_(0, rowt[r], recv[r, k, :len(per[r])])

    def _sync_engine_to_torch(self):
        # engine and torch share the stream on GPUs (Engine.set_stream); on the CPU simulator calls are synchronous
        pass

    def relinearize(self, c0, c1, c2, S, evk_a, evk_b, dig_polys=None):
        """3-part (1, s, s^2) ciphertexts over ctxt primes S -> 2-part over S | special (rows owned by this
        rank only).  c0,c1,c2: lists of Poly (batch items) holding at least the owned rows; c2 is consumed."""
        E = self.E
        Sp = sorted(set(S) | set(self.special))
        oS, oSp, oSpec = self.owned(S), self.owned(Sp), self.owned(self.special)
        nit = len(c0)
        # digits (src/DoubleCRT.cpp:479-561)
        remaining, nd = set(S), 0
        while remaining:
            remaining -= set(self.digits[nd])
            nd += 1
        dsets = [[i for i in S if i in self.digits[d]] for d in range(nd)]
        if dig_polys is None:
            dig_polys = [[E.poly() for _ in range(nd)] for _ in range(nit)]
        fused = E.N % 512 == 0 and nd <= 4       # streaming inner-product kernel: aliased own rows, folded scale / zero
        if not fused:
            # parts with handle 1 / base s: addPrimesAndScale(special) on the owned rows (src/Ctxt.cpp:764-768)
            if oS:
                E.scale_by_primes(c0, oS, self.special)
                E.scale_by_primes(c1, oS, self.special)
            if oSpec:
                E.zero_rows(c0, oSpec)
                E.zero_rows(c1, oSpec)
            for d in range(nd):
                od = self.owned(dsets[d])
                if od:
                    E.pointwise("copy", [dp[d] for dp in dig_polys], c2, od)
        for d in range(nd):
            col = [dp[d] for dp in dig_polys]
            if not dsets[d]:
                # an index set with a hole: no live prime in this digit.  The reference carries the zero polynomial for it and
                # still divides the later digits by the digit's full product (src/DoubleCRT.cpp:488-493,509-561)
                if oSp:
                    E.zero_rows(col, oSp)
                for j in range(d + 1, nd):
                    oj = self.owned(dsets[j])
                    if oj:
                        E.scale_by_primes(c2 if fused else [dp[j] for dp in dig_polys], oj, self.digits[d], inv=True)
                continue
            ys = [self._ybuf(("dig", it)) for it in range(nit)]
            # fused: digit d's own rows ARE c2's rows by now (the mixed-radix steps update c2 in place)
            self._exchange(c2 if fused else col, dsets[d], ys)
            tgt = self.owned([i for i in Sp if i not in dsets[d]])
            E.conv_from_y([y[1] for y in ys], dsets[d], tgt, 1, col, 0)
            for j in range(d + 1, nd):   # digits[j] -= digits[d]; digits[j] /= prod(full digit d)
                oj = self.owned(dsets[j])
                if oj:
                    E.sub_div_by_primes(c2 if fused else [dp[j] for dp in dig_polys], col, oj, self.digits[d])
        # evk inner product on the owned rows (src/Ctxt.cpp:191-230)
        if oSp:
            digs = [dp[:nd] for dp in dig_polys]
            if fused:
                P = 1
                for i in self.special:
                    P *= E.primes[i]
                inS = set(S)
                scal = [P % E.primes[r] if r in inS else 0 for r in oSp]
                own_dig = [next(d for d in range(nd) if r in dsets[d]) if r in inS else -1 for r in oSp]
                E.keyswitch_digits_fused(digs, oSp, evk_a[:nd], evk_b[:nd], c0, c1, scal, own=c2, own_dig=own_dig)
            else:
                E.keyswitch_digits(digs, oSp, evk_a[:nd], evk_b[:nd], c0, c1)
        return Sp

    def mod_down(self, parts, cur, keep, ptxt_space=1):
        """Ctxt::modDownToSet on row-sharded parts: drop cur\\keep (e.g. the special primes)."""
        E = self.E
        drop = [i for i in cur if i not in keep]
        if not drop:
            return
        ys = [self._ybuf(("md", k)) for k in range(len(parts))]
        self._exchange(parts, drop, ys)
        E.conv_from_y([y[1] for y in ys], drop, self.owned(keep), ptxt_space, parts, 1)
