from .genome import Genome, codes_of_string  # noqa: F401
