"""Stand-in for ``pyfaidx`` (absent from the image, no network) used ONLY by the reference arm of bench.py and
by tests/golden/make_golden.py.  It provides exactly what selene_sdk touches (genome.py:290-291,345-352,
model_predict.py:326,491-522): ``Fasta(path)`` with ``keys() / in / [chrom] / iteration / close()``; a record
supports ``len()``, ``str()``, ``.name`` and slicing to an object with ``.seq``, ``.reverse``, ``.complement``
(case-preserving, non-ACGT characters unchanged — pyfaidx's behaviour).  No arithmetic of the path lives here."""
_COMP = str.maketrans("ACGTacgt", "TGCAtgca")


class _Seq(object):
    __slots__ = ("seq", "name")

    def __init__(self, s, name=None):
        self.seq = s
        self.name = name

    def __len__(self):
        return len(self.seq)

    def __str__(self):
        return self.seq

    def __getitem__(self, sl):
        return _Seq(self.seq[sl], self.name)

    @property
    def reverse(self):
        return _Seq(self.seq[::-1], self.name)

    @property
    def complement(self):
        return _Seq(self.seq.translate(_COMP), self.name)


class Fasta(object):
    def __init__(self, path, **kw):
        self._d = {}
        with open(path, "rb") as fh:
            blob = fh.read()
        for rec in blob.split(b">")[1:]:
            nl = rec.find(b"\n")
            head = rec[:nl].decode("latin-1").strip()
            name = head.split()[0] if head else ""
            self._d[name] = _Seq(rec[nl + 1:].replace(b"\n", b"").replace(b"\r", b"").decode("latin-1"), name)

    def keys(self):
        return self._d.keys()

    def __contains__(self, k):
        return k in self._d

    def __getitem__(self, k):
        return self._d[k]

    def __iter__(self):
        return iter(self._d.values())

    def close(self):
        pass
