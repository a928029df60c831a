from .mesh import Mesh
from .mesh_cell import MeshCell
from .generators import box_tet_mesh, fractured_box_tet_mesh, square_tri_mesh
