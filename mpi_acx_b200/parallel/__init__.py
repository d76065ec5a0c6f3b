from .pipeline import PartitionedLink
from .ring import allpairs_exchange, ring_exchange, ring_exchange_prepost, ring_neighbours

__all__ = ["ring_exchange", "ring_exchange_prepost", "ring_neighbours", "allpairs_exchange", "PartitionedLink"]
