"""Minimal FASTA/FASTQ(.gz) reader for the tests (ids = whole header line, newlines stripped from sequences)."""
import gzip


def read_fastx(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        data = f.read()
    recs = []
    if not data:
        return recs
    lines = data.split(b"\n")
    if data[:1] == b">":
        name, chunks = None, []
        for ln in lines:
            ln = ln.rstrip(b"\r")
            if ln.startswith(b">"):
                if name is not None:
                    recs.append((name, b"".join(chunks)))
                name, chunks = ln[1:].decode(), []
            else:
                chunks.append(ln)
        if name is not None:
            recs.append((name, b"".join(chunks)))
    elif data[:1] == b"@":
        i = 0
        while i + 3 < len(lines) + 1 and i < len(lines) and lines[i].startswith(b"@"):
            recs.append((lines[i][1:].rstrip(b"\r").decode(), lines[i + 1].rstrip(b"\r")))
            i += 4
    return recs
