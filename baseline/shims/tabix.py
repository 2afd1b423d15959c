"""Stand-in for ``pytabix`` (absent from the image): ``open(path).query(chrom, start, end)`` returns the rows
(lists of str) with ``row_start < end and row_end > start`` — tabix's half-open overlap — and raises
``TabixError`` for an unknown contig (selene_sdk/targets/genomic_features.py:341-348)."""
import gzip


class TabixError(Exception):
    pass


class _Table(object):
    def __init__(self, path):
        self.by = {}
        op = gzip.open if path.endswith(".gz") else open
        with op(path, "rt") as fh:
            for line in fh:
                c = line.rstrip("\n").split("\t")
                if len(c) < 3:
                    continue
                self.by.setdefault(c[0], []).append(c)

    def query(self, chrom, start, end):
        if chrom not in self.by:
            raise TabixError("unknown contig")
        return [r for r in self.by[chrom] if int(r[1]) < end and int(r[2]) > start]


def open(path):  # noqa: A001  (pytabix's name)
    return _Table(path)
