"""Host-side container for a contig's fragments in the CSR form that crosses the C-ABI (fb_frags).

Mirrors what the reference keeps per read (src/types_structs.rs:68-85: seq_dict, qual_dict, positions,
first_position, last_position, counter_id) as flat arrays, and reads/writes the H-PoP text format the
reference itself parses (src/file_reader.rs:37-109) and writes (src/file_writer.rs:665-696), so a Rust
harness can consume identical inputs.
"""
import ctypes as C

import numpy as np

from ._cdefs import FbFrags, ptr, u8p, u32p, u64p


class Frags:
    def __init__(self, row_ptr, pos, allele, qual, first=None, last=None):
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.pos = np.ascontiguousarray(pos, dtype=np.uint32)
        self.allele = np.ascontiguousarray(allele, dtype=np.uint8)
        self.qual = np.ascontiguousarray(qual, dtype=np.uint8)
        self.n_reads = len(self.row_ptr) - 1
        self.nnz = int(self.row_ptr[-1]) if self.n_reads >= 0 and len(self.row_ptr) else 0
        if first is None:
            first = self.pos[self.row_ptr[:-1].astype(np.int64)] if self.n_reads else np.zeros(0, np.uint32)
        if last is None:
            last = self.pos[self.row_ptr[1:].astype(np.int64) - 1] if self.n_reads else np.zeros(0, np.uint32)
        self.first = np.ascontiguousarray(first, dtype=np.uint32)
        self.last = np.ascontiguousarray(last, dtype=np.uint32)

    # ---- C view --------------------------------------------------------------------------------
    def as_struct(self):
        s = FbFrags()
        s.n_reads = self.n_reads
        s.nnz = self.nnz
        s.row_ptr = ptr(self.row_ptr, u64p)
        s.first = ptr(self.first, u32p)
        s.last = ptr(self.last, u32p)
        s.pos = ptr(self.pos, u32p)
        s.allele = ptr(self.allele, u8p)
        s.qual = ptr(self.qual, u8p)
        s._keep = self  # keep the arrays alive as long as the struct is
        return s

    def read(self, i):
        a, b = int(self.row_ptr[i]), int(self.row_ptr[i + 1])
        return self.pos[a:b], self.allele[a:b], self.qual[a:b]

    # ---- ordering (types_structs.rs:87-93 Frag::cmp; floria.rs:289-293) ---------------------------
    def is_sorted(self):
        f = self.first.astype(np.int64)
        l = self.last.astype(np.int64)
        if self.n_reads < 2:
            return True
        ok = (f[1:] > f[:-1]) | ((f[1:] == f[:-1]) & (l[1:] <= l[:-1]))
        return bool(ok.all())

    @staticmethod
    def from_reads(reads, sort=True):
        """reads: list of (pos[], allele[], qual[]) with pos ascending. Sorted by Frag::cmp when sort=True."""
        firsts = np.array([r[0][0] for r in reads], dtype=np.int64)
        lasts = np.array([r[0][-1] for r in reads], dtype=np.int64)
        order = np.lexsort((np.arange(len(reads)), -lasts, firsts)) if sort else np.arange(len(reads))
        row_ptr = np.zeros(len(reads) + 1, dtype=np.uint64)
        pos, al, ql = [], [], []
        for k, i in enumerate(order):
            p, a, q = reads[i]
            row_ptr[k + 1] = row_ptr[k] + np.uint64(len(p))
            pos.append(np.asarray(p, dtype=np.uint32))
            al.append(np.asarray(a, dtype=np.uint8))
            ql.append(np.asarray(q, dtype=np.uint8))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        fr = Frags(row_ptr, cat(pos, np.uint32), cat(al, np.uint8), cat(ql, np.uint8))
        fr.order = order
        return fr

    def subset(self, ids):
        """New Frags made of the given reads (in the given order), counter_ids renumbered."""
        ids = np.asarray(ids, dtype=np.int64)
        lens = (self.row_ptr[1:] - self.row_ptr[:-1]).astype(np.int64)[ids]
        row_ptr = np.zeros(len(ids) + 1, dtype=np.uint64)
        row_ptr[1:] = np.cumsum(lens)
        starts = self.row_ptr[:-1].astype(np.int64)[ids]
        idx = np.repeat(starts - row_ptr[:-1].astype(np.int64), lens) + np.arange(int(row_ptr[-1]))
        return Frags(row_ptr, self.pos[idx], self.allele[idx], self.qual[idx], self.first[ids], self.last[ids])

    # ---- H-PoP frags text format -------------------------------------------------------------------
    def write_hpop(self, path, ids=None):
        """file_writer.rs:665-696 write_frags_file: `nblocks\\tid\\t(start\\talleles\\t)*quals`."""
        with open(path, "w", encoding="latin-1") as fh:
            for i in range(self.n_reads):
                p, a, q = self.read(i)
                blocks = []
                s = 0
                for k in range(1, len(p) + 1):
                    if k == len(p) or p[k] != p[k - 1] + 1:
                        blocks.append((int(p[s]), "".join(str(int(x)) for x in a[s:k])))
                        s = k
                name = ids[i] if ids is not None else f"read{i}"
                quals = "".join(chr(int(x) + 33) if int(x) + 33 <= 255 else chr(int(x)) for x in q)
                fh.write(f"{len(blocks)}\t{name}\t" + "".join(f"{b}\t{s_}\t" for b, s_ in blocks) + quals + "\n")

    @staticmethod
    def read_hpop(path):
        """file_reader.rs:37-109 get_frags_container."""
        reads, names = [], []
        with open(path, "r", encoding="latin-1") as fh:
            for line in fh:
                v = line.rstrip("\n").split("\t")
                nb = int(v[0])
                names.append(v[1])
                pos, al = [], []
                for b in range(nb):
                    start = int(v[2 * b + 2])
                    for j, c in enumerate(v[2 * b + 3]):
                        pos.append(start + j)
                        al.append(int(c))
                q = [ord(c) - 33 for c in v[-1]][: len(pos)]
                reads.append((pos, al, q))
        fr = Frags.from_reads(reads, sort=False)
        fr.names = names
        return fr
