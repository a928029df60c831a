"""``basix.cell.sub_entity_connectivity``: only slot 0 (the vertices of each sub-entity) is filled -- the one
R/src/basis/parametric_transformation.py:12-13, 34-35, 61-64 reads."""
from oracle.ref_basis import SUB


def sub_entity_connectivity(celltype):
    dim = celltype.value
    return [[[list(v), [], [], []] for v in SUB[dim][d]] for d in range(dim + 1)]
