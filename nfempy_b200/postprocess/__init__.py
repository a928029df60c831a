from .l2_error_post_processor import div_error, l2_error
from .projectors import l2_projector
