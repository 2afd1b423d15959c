from .genomic_features import GenomicFeatures  # noqa: F401
