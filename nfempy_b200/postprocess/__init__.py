from .l2_error_post_processor import div_error, div_scaled_error, l2_error, l2_error_projected
from .projectors import l2_projector
from .solution_norms_post_processor import div_norm, l2_norm
