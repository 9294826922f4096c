"""The comparison baseline named in BASELINE.json: same placement and training, neighbour exchange over ``torch.distributed``
NCCL, aggregation with stock PyTorch ops — written the way a competent NCCL/PyTorch implementation would be (flat rows,
collectives where the topology allows them, vectorised tensor math, no per-key Python loops and no ``.item()`` syncs).

* full-mesh FedAvg is ONE ``all_reduce`` of the locally summed published rows;
* every other topology ships each needed flat row once per destination rank with ``batch_isend_irecv``
  (the exchange of reference ``murmura/distributed/node_process.py:227-276`` on NCCL instead of ZeroMQ);
* FedAvg / Krum / BALANCE / Sketchguard aggregate stacked rows with a handful of fused tensor ops per node; UBAR and
  EvidentialTrust (forward passes through foreign weights) reuse the parity aggregator classes on the received rows.

None of the hand-written exchange / aggregation kernels run on this path (``b200.transport: nccl``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from murmura_b200.aggregation.balance import decayed_factor


class NcclBaseline:
    def __init__(self, eng):
        self.eng = eng
        L = eng.layout
        self.row_len = L.Pf_pad + max(L.Pi, 0)                     # [float region | int buffers as float]
        self._classes: Dict[int, object] = {}
        self._sg_hist = None

    # ---- published copies (attack applied with torch ops) ---------------------------------------------------------------
    def _published(self) -> torch.Tensor:
        eng, L = self.eng, self.eng.layout
        pub = torch.zeros(max(eng.V, 1), self.row_len, device=eng.device)
        if eng.V == 0:
            return pub
        pub[: eng.V, : L.Pf_pad] = eng.live[: eng.V, : L.Pf_pad]
        if L.Pi:
            pub[: eng.V, L.Pf_pad:] = eng.ints[: eng.V, : L.Pi].float()
        spec = eng.attack_spec
        if spec is not None:
            byz = [vn.slot for vn in eng.nodes if vn.byzantine]
            if byz:
                idx = torch.tensor(byz, device=eng.device)
                rows = pub[idx, : L.Pf]
                rows = rows * float(spec["scale"])
                if spec["noise_std"]:
                    rows = rows + torch.randn_like(rows) * float(spec["noise_std"])
                pub[idx, : L.Pf] = rows
        elif eng.attack is not None:
            for vn in eng.nodes:
                if vn.byzantine:
                    state = eng.attack.apply_attack(node_id=vn.gid, model_state=dict(L.row_views(eng.live[vn.slot].clone(), None)), round_num=eng.round_idx)
                    fresh = torch.zeros(L.stride, device=eng.device)
                    for k, v in L.row_views(fresh, None).items():
                        v.copy_(state[k])
                    pub[vn.slot, : L.Pf_pad] = fresh[: L.Pf_pad]
        return pub

    # ---- exchange -------------------------------------------------------------------------------------------------------------
    def _exchange(self, pub: torch.Tensor, neighbors: List[List[int]]) -> Dict[int, torch.Tensor]:
        """Rows of every node some local node listens to (gid → flat row)."""
        eng, pl = self.eng, self.eng.placement
        rows: Dict[int, torch.Tensor] = {vn.gid: pub[vn.slot] for vn in eng.nodes}
        if eng.world == 1:
            return rows
        import torch.distributed as dist
        need: Dict[Tuple[int, int], None] = {}                     # (src gid, dst rank), deterministic order on every rank
        for i in range(eng.N):
            for j in neighbors[i]:
                if pl.rank_of[i] != pl.rank_of[j]:
                    need[(j, int(pl.rank_of[i]))] = None
        ops = []
        for src, rd in need:
            rs = int(pl.rank_of[src])
            if rs == eng.rank:
                ops.append(dist.P2POp(dist.isend, pub[int(pl.slot_of[src])], rd))
            elif rd == eng.rank:
                buf = torch.empty(self.row_len, device=eng.device)
                rows[src] = buf
                ops.append(dist.P2POp(dist.irecv, buf, rs))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return rows

    # ---- write-back --------------------------------------------------------------------------------------------------------------
    def _store(self, vn, out: torch.Tensor, ints_from: torch.Tensor = None) -> None:
        eng, L = self.eng, self.eng.layout
        eng.live[vn.slot, : L.Pf].copy_(out[: L.Pf])
        if L.Pi and ints_from is not None:
            eng.ints[vn.slot, : L.Pi].copy_(ints_from[L.Pf_pad: L.Pf_pad + L.Pi].round().long())

    # ---- aggregation families ----------------------------------------------------------------------------------------------------
    def aggregate(self, neighbors: List[List[int]]) -> None:
        eng, L = self.eng, self.eng.layout
        fam = eng.family
        pub = self._published()
        full_mesh = all(len(set(neighbors[i]) - {i}) == eng.N - 1 for i in range(eng.N))
        if fam == "fedavg" and full_mesh and not eng.opt.fault_drop_edges:
            total = pub[: eng.V, : L.Pf].sum(0) if eng.V else torch.zeros(L.Pf, device=eng.device)
            if eng.world > 1:
                import torch.distributed as dist
                dist.all_reduce(total)
            # the reference averages [own] + neighbours where own is the UNATTACKED state: swap the local published copy for it
            for vn in eng.nodes:
                mean = (total - pub[vn.slot, : L.Pf] + eng.live[vn.slot, : L.Pf]) / eng.N
                eng.live[vn.slot, : L.Pf].copy_(mean)
            return
        rows = self._exchange(pub, neighbors)
        a = eng.aggregator
        outs = []
        for vn in eng.nodes:
            own = torch.cat([eng.live[vn.slot, : L.Pf_pad], eng.ints[vn.slot, : L.Pi].float()]) if L.Pi else eng.live[vn.slot, : L.Pf_pad].clone()
            nbr = [rows[j] for j in neighbors[vn.gid] if j != vn.gid]
            if not nbr:
                outs.append((vn, None, None)); continue
            X = torch.stack(nbr)                                   # [d, row]
            if fam == "fedavg":
                out = (own[: L.Pf] + X[:, : L.Pf].sum(0)) / (len(nbr) + 1)
                outs.append((vn, out, None))
            elif fam == "krum":
                m, c = len(nbr) + 1, int(a.num_compromised)
                if c >= (m - 2) / 2:
                    outs.append((vn, None, None)); continue
                A = torch.cat([own[None, : L.Pf], X[:, : L.Pf]])   # float tensors only (reference base.py:118-135)
                D = torch.stack([(A - A[i]).norm(dim=1) for i in range(m)])
                k = max(1, m - c - 2)
                D = D + torch.diag(torch.full((m,), float("inf"), device=D.device))
                score = D.sort(dim=1).values[:, :k].sum(1)
                win = score.argmin()
                full = torch.cat([own[None], X]).index_select(0, win.view(1))[0]
                outs.append((vn, full[: L.Pf], full))
            elif fam == "balance":
                d = (X - own).norm(dim=1)                          # all keys incl. int buffers (reference balance.py:91-106)
                thr = decayed_factor(a.gamma, a.kappa, eng.round_idx, a.total_rounds) * own.norm()
                acc = d <= thr
                closest = torch.zeros_like(acc); closest[d.argmin()] = True
                acc = torch.where(acc.sum() >= a.min_neighbors, acc, acc | closest)
                w = acc.float() / acc.float().sum().clamp_min(1.0)
                full = a.alpha * own + (1 - a.alpha) * (w[:, None] * X).sum(0)
                outs.append((vn, full[: L.Pf], full))
            elif fam == "sketchguard":
                full = self._sketchguard(vn, own, X)
                outs.append((vn, full[: L.Pf], full))
            else:
                return self._classes_path(neighbors, rows)
        for vn, out, full in outs:
            if out is not None:
                self._store(vn, out, full)

    def _sketch(self, v: torch.Tensor) -> torch.Tensor:
        eng = self.eng
        if not hasattr(self, "_bucket"):
            t = eng.sk_table.view(torch.int16).to(torch.int32) & 0xFFFF
            self._bucket = (t & 0x7FFF).long()[: eng.layout.Pf]
            self._sign = torch.where((t >> 15) > 0, -1.0, 1.0)[: eng.layout.Pf].to(eng.device)
        return torch.zeros(v.shape[0], int(eng.aggregator.sketch_size), device=v.device).index_add_(1, self._bucket, v[:, : eng.layout.Pf] * self._sign)

    def _sketchguard(self, vn, own: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
        eng, a = self.eng, self.eng.aggregator
        if self._sg_hist is None:
            self._sg_hist = torch.ones(max(eng.V, 1), 3, device=eng.device)            # last three acceptance rates (1 = no history)
        s = self._sketch(torch.cat([own[None], X]))
        d = (s[1:] - s[0]).norm(dim=1)
        hist = self._sg_hist[vn.slot]
        factor = torch.where(hist.mean() < 0.3, 1.5, 1.0)
        thr = decayed_factor(a.gamma, a.kappa, eng.round_idx, a.total_rounds) * factor * s[0].norm()
        acc = d <= thr
        self._sg_hist[vn.slot] = torch.cat([hist[1:], acc.float().mean().view(1)])
        closest = torch.zeros_like(acc); closest[d.argmin()] = True
        acc = torch.where(acc.sum() >= a.min_neighbors, acc, acc | closest)
        w = acc.float() / acc.float().sum().clamp_min(1.0)
        full = a.alpha * own + (1 - a.alpha) * (w[:, None] * X).sum(0)
        if eng.layout.Pi:                                                                # int buffers: α·own + (1−α)·first accepted
            first = X[acc.float().argmax()]
            full[eng.layout.Pf_pad:] = a.alpha * own[eng.layout.Pf_pad:] + (1 - a.alpha) * first[eng.layout.Pf_pad:]
        return full

    # ---- forward-based filters: the parity classes on the received rows --------------------------------------------------------
    def _classes_path(self, neighbors: List[List[int]], rows: Dict[int, torch.Tensor]) -> None:
        from murmura_b200.parallel.engine import copy_aggregator
        eng, L = self.eng, self.eng.layout
        results = []
        for vn in eng.nodes:
            agg = self._classes.setdefault(vn.gid, copy_aggregator(eng.aggregator))
            own = {k: v.clone() for k, v in L.row_views(eng.live[vn.slot], eng.ints[vn.slot]).items()}
            nbrs = {}
            for j in neighbors[vn.gid]:
                if j == vn.gid:
                    continue
                r = torch.zeros(L.stride, device=eng.device); r[: L.Pf_pad] = rows[j][: L.Pf_pad]
                ints = rows[j][L.Pf_pad: L.Pf_pad + L.Pi].round().long() if L.Pi else None
                nbrs[j] = dict(L.row_views(r, ints))
            loader = [(eng._inputs(vn, vn.X[: max(vn.eb, 100)]), vn.y[: max(vn.eb, 100)])]
            results.append(agg.aggregate(node_id=vn.gid, own_state=own, neighbor_states=nbrs, round_num=eng.round_idx, train_loader=loader,
                                         model_template=vn.model, device=eng.device))
        for vn, st in zip(eng.nodes, results):
            views = L.row_views(eng.live[vn.slot], eng.ints[vn.slot])
            for k, v in st.items():
                views[k].copy_(v.to(views[k].dtype))
