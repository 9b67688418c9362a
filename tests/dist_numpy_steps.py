"""TEST INFRASTRUCTURE: a numpy + oracle implementation of the per-rank "steps"
interface of raven_b200/distributed.py, so that the exchange schedule (counts,
all-to-all, all-reduce, all-gather, partition bounds, flush schedule) runs on
CPU tensors over gloo and can be compared with the single-process oracle.
Each step restates what the matching rvn_dist_* entry point does on the GPU."""
import numpy as np
import torch

MAXOCC = 0xFFFFFFFF


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a).view(dtype).copy())


class NumpySteps:
    def __init__(self, oracle, reads, k=15, w=5):
        self.o = oracle
        self.rs = reads
        self.oreads = oracle.reads(reads)
        self.eng = oracle.engine(k, w)
        self.lens = np.asarray(reads.lens, dtype=np.uint32)

    # -- sketch of reads [first,last), stably split by value mod parts
    def sketch_split(self, first, last, parts, minhash):
        if last > first:
            s = self.o.sketch(self.eng, self.oreads, first, last, bool(minhash))
            val, org = s["value"], s["origin"]
        else:
            val = org = np.zeros(0, np.uint64)
        owner = (val % np.uint64(parts)).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        cnt = np.bincount(owner, minlength=parts).tolist()
        return _t(val[order], np.int64), _t(org[order], np.int64), cnt

    def max_threshold(self, first, last):
        if last <= first:
            return 0
        s = self.o.sketch(self.eng, self.oreads, first, last, True)
        return int(s["value"].max()) if s["value"].size else 0

    def build_index(self, val, org, bases, limit=None):
        # (the tiers are an implementation detail of the device index: same results)
        v = val.numpy().view(np.uint64)
        o = org.numpy().view(np.uint64)
        order = np.argsort(v, kind="stable")
        self.i_val, self.i_org = v[order], o[order]
        self.occ = MAXOCC

    def histogram(self):
        if self.i_val.size == 0:
            return torch.zeros(65536, dtype=torch.int64), 0
        _, counts = np.unique(self.i_val, return_counts=True)
        h = np.bincount(np.minimum(counts, 65535), minlength=65536).astype(np.int64)
        return torch.from_numpy(h), int(counts.size)

    def set_occurrence(self, hist, n_keys, frequency):
        if frequency == 0 or n_keys == 0:
            self.occ = MAXOCC
            return self.occ
        rank = min(int((1 - frequency) * float(n_keys)), n_keys - 1)
        cum = np.cumsum(hist[:65535].astype(np.uint64))
        length = int(np.searchsorted(cum, rank, side="right"))
        assert length < 65535
        self.occ = length + 1
        return self.occ

    # -- hits of the query records against this slice, runs by owner of the read
    def hits_split(self, qval, qorg, parts, n_query):
        qv = qval.numpy().view(np.uint64)
        qo = qorg.numpy().view(np.uint64)
        assert np.all(np.diff((qo >> np.uint64(32)).astype(np.int64)) >= 0)  # the contract
        lo = np.searchsorted(self.i_val, qv, side="left")
        hi = np.searchsorted(self.i_val, qv, side="right")
        n = hi - lo
        n[n > self.occ] = 0
        qi = np.repeat(np.arange(qv.size), n)
        pi = np.concatenate([np.arange(a, a + c) for a, c in zip(lo, n)]) if qi.size \
            else np.zeros(0, np.int64)
        lorg, rorg = qo[qi], self.i_org[pi]
        lid, rid = lorg >> np.uint64(32), rorg >> np.uint64(32)
        keep = (lid != rid) & ~(lid > rid)
        lorg, rorg, lid, rid = lorg[keep], rorg[keep], lid[keep], rid[keep]
        lpos = (lorg & np.uint64(0xFFFFFFFF)) >> np.uint64(1)
        rpos = (rorg & np.uint64(0xFFFFFFFF)) >> np.uint64(1)
        strand = ((lorg & np.uint64(1)) == (rorg & np.uint64(1))).astype(np.uint64)
        diag = np.where(strand == 0, rpos + lpos, rpos - lpos + np.uint64(3 << 30))
        grp = (((rid << np.uint64(1)) | strand) << np.uint64(32)) | diag
        pos = (lpos << np.uint64(32)) | rpos
        dest = (lid % np.uint64(parts)).astype(np.int64)
        order = np.argsort(dest, kind="stable")
        cnt = np.bincount(dest, minlength=parts).tolist()
        return (_t(grp[order], np.int64), _t(pos[order], np.int64),
                _t(lid[order].astype(np.uint32), np.int32), cnt)

    @staticmethod
    def _check_runs(keys, run_counts):
        assert sum(run_counts) == keys.size
        at = 0
        for c in run_counts:
            assert np.all(np.diff(keys[at:at + c].astype(np.int64)) >= 0)
            at += c

    def chain(self, grp, pos, lhs, run_counts, parts, rank, n_query):
        g = grp.numpy().view(np.uint64)
        p = pos.numpy().view(np.uint64)
        l = lhs.numpy().view(np.uint32)
        self._check_runs(l, run_counts)  # the contract of rvn_dist_chain
        assert np.all(l % parts == rank) and (l.size == 0 or l.max() < n_query)
        order = np.argsort(l, kind="stable")
        g, p, l = g[order], p[order], l[order]
        out = []
        for r in range(rank, n_query, parts):
            a, b = np.searchsorted(l, [r, r + 1])
            if b > a:
                out.append(self.o.chain(self.eng, r, g[a:b], p[a:b]))
        self.ovl = np.concatenate(out) if out else np.zeros((0, 8), np.uint32)
        self.mapped += self.ovl.shape[0]
        return self.ovl.shape[0]

    def overlaps_split(self, parts, rank):
        runs = [self.ovl if d == rank else self.ovl[self.ovl[:, 3] % parts == d]
                for d in range(parts)]
        ov = np.concatenate(runs)
        return (torch.from_numpy(ov.view(np.int32).reshape(-1, 8).copy()),
                [r.shape[0] for r in runs])

    # -- construct.cc:66-112 for the owned reads
    def stage1_begin(self, parts, rank):
        self.parts, self.rank = parts, rank
        self.own = list(range(rank, self.rs.n, parts))
        self.piles = {r: np.zeros(int(self.lens[r]) >> 4, np.uint16) for r in self.own}
        self.lists = {r: np.zeros((0, 8), np.uint32) for r in self.own}
        self.mapped = 0

    def stage1_add(self, ovl, run_counts, n_query, kmax, qb):
        qb = qb or (1 << 30)
        ovl = ovl.numpy().view(np.uint32).reshape(-1, 8)
        self._check_runs(ovl[:, 0], run_counts)
        ovl = ovl[np.argsort(ovl[:, 0], kind="stable")]  # merge by query
        off = np.searchsorted(ovl[:, 0], np.arange(n_query + 1))
        bases, k0 = 0, 0
        for k in range(n_query):
            bases += int(self.lens[k])
            if k != n_query - 1 and bases < qb:
                continue
            bases = 0
            new = {}
            for o in ovl[int(off[k0]):int(off[k + 1])]:
                if o[0] % self.parts == self.rank:
                    new.setdefault(int(o[0]), []).append(o)
                if o[3] % self.parts == self.rank:
                    new.setdefault(int(o[3]), []).append(o[[3, 4, 5, 0, 1, 2, 6, 7]])
            for r, add in new.items():
                add = np.array(add, dtype=np.uint32).reshape(-1, 8)
                self.piles[r] = self.o.pile_add_layers(r, self.piles[r], add)
                self.lists[r] = self.o.truncate(np.concatenate([self.lists[r], add]), kmax)
            k0 = k + 1

    def stage1_end(self):
        pass

    def stage1_results(self):
        n = len(self.own)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum([self.lists[r].shape[0] for r in self.own])
        poff = np.zeros(n + 1, np.uint64)
        poff[1:] = np.cumsum([self.piles[r].size for r in self.own])
        ov = np.concatenate([self.lists[r] for r in self.own]) if n else np.zeros((0, 8))
        pl = np.concatenate([self.piles[r] for r in self.own]) if n else np.zeros(0)
        return dict(overlaps=ov.reshape(-1, 8).astype(np.uint32), ovl_off=off,
                    pile=pl.astype(np.uint16), pile_off=poff, num_mapped=self.mapped)
