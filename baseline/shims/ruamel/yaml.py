"""Stub for ``ruamel.yaml`` (absent from the image; only selene_sdk/utils/config_utils.py's CLI path uses it)."""


class YAML(object):
    def __init__(self, *a, **kw):
        raise ImportError("ruamel.yaml is not installed in this image")
