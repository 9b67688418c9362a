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
    def sketch_split(self, first, last, parts):
        if last > first:
            s = self.o.sketch(self.eng, self.oreads, first, last, True)
            val, org = s["value"], s["origin"]
        else:
            val = org = np.zeros(0, np.uint64)
        owner = (val % np.uint64(parts)).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        cnt = np.bincount(owner, minlength=parts).tolist()
        return _t(val[order], np.int64), _t(org[order], np.int64), cnt

    def build_index(self, val, org, bases):
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

    # -- hits of the query records against this slice, split by read owner
    def hits_split(self, qval, qorg, bounds):
        qv = qval.numpy().view(np.uint64)
        qo = qorg.numpy().view(np.uint64)
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
        b = np.asarray(bounds, dtype=np.uint64)
        dest = np.searchsorted(b[1:-1], lid, side="right")
        order = np.argsort(dest, kind="stable")
        cnt = np.bincount(dest, minlength=len(bounds) - 1).tolist()
        return (_t(grp[order], np.int64), _t(pos[order], np.int64),
                _t(lid[order].astype(np.uint32), np.int32), cnt)

    def chain(self, grp, pos, lhs, first, last):
        g = grp.numpy().view(np.uint64)
        p = pos.numpy().view(np.uint64)
        l = lhs.numpy().view(np.uint32)
        out, cnt = [], []
        order = np.argsort(l, kind="stable")
        g, p, l = g[order], p[order], l[order]
        starts = np.searchsorted(l, np.arange(first, last + 1))
        for r in range(first, last):
            a, b = starts[r - first], starts[r - first + 1]
            ov = self.o.chain(self.eng, r, g[a:b], p[a:b]) if b > a else \
                np.zeros((0, 8), np.uint32)
            out.append(ov)
            cnt.append(ov.shape[0])
        ov = np.concatenate(out) if out else np.zeros((0, 8), np.uint32)
        return (torch.from_numpy(ov.view(np.int32).reshape(-1, 8).copy()),
                torch.tensor(cnt, dtype=torch.int32))

    # -- construct.cc:66-112 on the complete ordered overlap list
    def stage1_begin(self):
        n = self.rs.n
        self.piles = [np.zeros(int(x) >> 4, np.uint16) for x in self.lens]
        self.lists = [np.zeros((0, 8), np.uint32) for _ in range(n)]
        self.mapped = 0

    def stage1_add(self, ovl, off, n_query, kmax, qb):
        qb = qb or (1 << 30)
        ovl = ovl.numpy().view(np.uint32).reshape(-1, 8)
        n = self.rs.n
        seen = [x.shape[0] for x in self.lists]
        bases, k0 = 0, 0
        for k in range(n_query):
            bases += int(self.lens[k])
            if k != n_query - 1 and bases < qb:
                continue
            bases = 0
            new = [[] for _ in range(n)]
            for o in ovl[int(off[k0]):int(off[k + 1])]:
                self.mapped += 1
                new[o[0]].append(o)
                new[o[3]].append(o[[3, 4, 5, 0, 1, 2, 6, 7]])
            for r in range(n):
                if not new[r]:
                    continue
                add = np.array(new[r], dtype=np.uint32).reshape(-1, 8)
                self.piles[r] = self.o.pile_add_layers(r, self.piles[r], add)
                self.lists[r] = self.o.truncate(np.concatenate([self.lists[r], add]), kmax)
            k0 = k + 1

    def stage1_end(self):
        pass

    def stage1_results(self, n):
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum([x.shape[0] for x in self.lists])
        poff = np.zeros(n + 1, np.uint64)
        poff[1:] = np.cumsum([x.size for x in self.piles])
        return dict(overlaps=np.concatenate(self.lists).reshape(-1, 8), ovl_off=off,
                    pile=np.concatenate(self.piles), pile_off=poff, num_mapped=self.mapped)
