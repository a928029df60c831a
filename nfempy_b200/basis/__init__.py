from .quadrature import make_quadrature
from .reference_element import ReferenceElement, create_element
