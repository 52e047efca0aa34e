"""Independent Python decoder of skani v0.3.0's on-disk sketch formats (bincode 1.3 defaults applied to
src/params.rs:137-146, src/types.rs:253-277, src/sketch_db.rs:10-15).  Test infrastructure: used to check what
skani_b200/cli/sketch_db.hpp writes, byte for byte, without going through that code."""
import struct

import numpy as np

DNA_TO_AA = b"KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"      # src/types.rs:27-28
AA_CODE = {b"A": 0, b"N": 2, b"D": 3, b"C": 4, b"E": 5, b"F": 6, b"G": 7, b"H": 8, b"I": 9, b"K": 10, b"L": 11, b"M": 12, b"P": 13,
           b"Q": 14, b"R": 15, b"S": 16, b"T": 17, b"V": 18, b"W": 19, b"Y": 20, b"*": 21}        # src/params.rs:150-180 (later 'R' wins)


class Cur:
    def __init__(self, b, o=0):
        self.b, self.o = b, o

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def string(self):
        n = self.take("Q")
        s = self.b[self.o:self.o + n]
        self.o += n
        return s.decode()

    def array(self, dtype, n):
        a = np.frombuffer(self.b, dtype, n, self.o)
        self.o += a.nbytes
        return a


def params(c):
    p = dict(c=c.take("Q"), k=c.take("Q"), marker_c=c.take("Q"), use_syncs=c.take("B"), use_aa=c.take("B"))
    n = c.take("Q")
    p["aa_encoding"] = c.array("<u8", n).tolist()
    n = c.take("Q")
    p["aa_letters"] = bytes(c.array("u1", n))
    p["orf_size"] = c.take("Q")
    return p


def expected_params_bytes(cc, k, m):
    out = struct.pack("<QQQBB", cc, k, m, 0, 0) + struct.pack("<Q", 64)
    out += b"".join(struct.pack("<Q", AA_CODE[bytes([x])]) for x in DNA_TO_AA)
    out += struct.pack("<Q", 64) + DNA_TO_AA + struct.pack("<Q", 30)
    assert len(out) == 626
    return out


def sketch(c):
    s = dict(file_name=c.string())
    tag = c.take("B")
    assert tag in (0, 1)
    s["has_seeds"] = bool(tag)
    recs = []
    entries = []
    if tag:
        n = c.take("Q")
        e = c.array(np.dtype([("k", "<u4"), ("v", "<u8")]), n)
        entries = list(zip(e["k"].tolist(), e["v"].tolist()))
    multi = []
    for _ in range(c.take("Q")):
        n = c.take("Q")
        multi.append(c.array(np.dtype([("pos", "<u4"), ("cc", "<u4")]), n))
    used = set()
    for key, v in entries:
        if v & 1:
            packed = v >> 1
            recs.append((key, packed >> 31, packed & 0x7FFFFFFF))
        else:
            assert (v >> 1) not in used
            used.add(v >> 1)
            for r in multi[v >> 1]:
                recs.append((key, int(r["pos"]), int(r["cc"])))
            assert len(multi[v >> 1]) >= 2
    assert len(used) == len(multi)
    s["n_keys"] = len(entries)
    s["records"] = sorted(recs, key=lambda r: (r[0], r[2] >> 1, r[1]))
    s["contigs"] = [c.string() for _ in range(c.take("Q"))]
    s["total_len"] = c.take("Q")
    s["contig_lengths"] = c.array("<u4", c.take("Q")).tolist()
    s["repetitive_kmers"] = c.take("Q")
    s["markers"] = sorted(c.array("<u8", c.take("Q")).tolist())
    s["marker_c"], s["c"], s["k"], s["contig_order"] = c.take("QQQQ")
    s["individual_contig"], s["amino_acid"] = c.take("BB")
    return s


def read_db(d):
    """-> (params, [sketch...] from sketches.db via index.db, [marker sketch...] from markers.bin, index entries)"""
    import os
    ix = Cur(open(os.path.join(d, "index.db"), "rb").read())
    index = []
    for _ in range(ix.take("Q")):
        index.append((ix.string(), ix.take("Q"), ix.take("Q")))
    assert ix.o == len(ix.b)
    db = open(os.path.join(d, "sketches.db"), "rb").read()
    sk, par, end = [], None, 0
    for name, off, ln in index:
        assert off == end                      # concatenated in arrival order, no gaps (src/sketch_db.rs:45-62)
        c = Cur(db, off)
        par = params(c)
        s = sketch(c)
        assert c.o == off + ln and s["file_name"] == name
        sk.append(s)
        end = off + ln
    assert end == len(db)
    mb = Cur(open(os.path.join(d, "markers.bin"), "rb").read())
    mpar = params(mb)
    mk = [sketch(mb) for _ in range(mb.take("Q"))]
    assert mb.o == len(mb.b)
    assert par is None or mpar == par
    return mpar, sk, mk, index
