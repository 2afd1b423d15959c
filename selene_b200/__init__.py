"""selene_b200 — B200-native (sm_100a) implementation of Selene's batched sequence -> prediction
hot path, behind the selene_sdk Python API.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
