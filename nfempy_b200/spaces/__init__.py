from .discrete_space import DiscreteSpace, DoFMap, MeshTopology
from .product_space import ProductSpace
