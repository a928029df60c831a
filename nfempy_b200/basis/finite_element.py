"""Per-cell element view mirroring the attribute surface of the reference ``FiniteElement``
(R/src/basis/finite_element.py:11-126): ``data.cell``, ``data.mesh``, ``data.dof.*``, ``data.bc_entities``.

The reference re-tabulates and re-maps the basis on the host for every cell and every call
(R/src/basis/finite_element.py:128-186).  In this package that work happens inside the CUDA kernels for
whole batches of cells, so the per-cell ``evaluate_mapping`` / ``evaluate_basis`` host methods do not exist:
calling them raises instead of silently computing on the CPU.
"""
import numpy as np


class _Dof:
    def __init__(self, ref):
        self.entity_dofs = ref.entity_dofs
        self.num_entity_dofs = ref.num_entity_dofs
        self.transformations_are_identity = ref.dof_transformations_are_identity
        self.n_dof = ref.n_phi
        self._ref = ref

    @property
    def transformations(self):
        return self._ref.entity_transformations()


class _Data:
    def __init__(self, space, cell_id, ref):
        self.mesh = space.mesh
        self.cell = space.mesh.cells[cell_id]
        self.dof = _Dof(ref)
        dim = self.cell.dimension
        if dim >= 1:
            facets = self.cell.sub_cells_ids[dim - 1]
            self.bc_entities = np.array([space.mesh.material[f] != -1 for f in facets])
        else:
            self.bc_entities = np.zeros(0, dtype=bool)


class FiniteElementView:
    def __init__(self, space, cell_id, ref_element, bc=False):
        self.generator_family = space.generator_family
        self.k_order = ref_element.degree if ref_element is not None else 0
        self.mesh = space.mesh
        self.discontinuous = space.discontinuous
        self.integration_order = space.integration_order
        self.basis_generator = ref_element
        self.data = _Data(space, cell_id, ref_element)

    @property
    def family(self):
        return self.basis_generator.family

    def evaluate_mapping(self, points, storage=False):
        raise NotImplementedError(
            "per-cell host mapping is not part of the GPU path; use nfempy_b200.geometry.evaluate_mapping_batch"
        )

    def evaluate_basis(self, points, jac, det_jac, inv_jac):
        raise NotImplementedError(
            "per-cell host tabulation is not part of the GPU path; basis tables are applied inside the kernels"
        )
