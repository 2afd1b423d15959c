from .random_positions_sampler import RandomPositionsSampler  # noqa: F401
