"""Stub for ``h5py`` (absent from the image): selene_sdk imports it at module scope (predict_handlers/handler.py:12);
the reference arm only uses TSV / in-memory outputs, so opening a file is an error."""


class File(object):
    def __init__(self, *a, **kw):
        raise ImportError("h5py is not installed in this image; the reference arm uses output_format='tsv'")
