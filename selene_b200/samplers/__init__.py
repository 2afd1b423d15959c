from .random_positions_sampler import RandomPositionsSampler  # noqa: F401
from .dataloader import H5DataLoader, H5Dataset, unpackbits_sequence, unpackbits_targets  # noqa: F401
