from .weak_form import WeakForm, BatchedWeakForm
from .le_primal_weak_form import (LEPrimalWeakForm, LEPrimalWeakFormBCDirichlet, LEPrimalWeakFormBCNeumann,
                                  LEPrimalWeakFormBCNormalDirichlet)
from .laplace_primal_weak_form import LaplacePrimalWeakForm, LaplacePrimalWeakFormBCDirichlet
from .laplace_dual_weak_form import LaplaceDualWeakForm, LaplaceDualWeakFormBCDirichlet
from .sundus_weak_form import SundusDualWeakForm, SundusDualWeakFormBCDirichlet
from .l2_projector_weak_form import L2ProjectorWeakForm
from .degenerate_elliptic_weak_form import DegenerateEllipticWeakForm, DegenerateEllipticWeakFormBCDirichlet
from .unscaled_elliptic_weak_form import UnscaledEllipticWeakForm, UnscaledEllipticWeakFormBCDirichlet
from .scaled_to_physical_l2_projection import ScaledToPhysicalL2Projection
from .tensor_mixed import (LEDualWeakForm, LEDualWeakFormBCDirichlet, LCEDualWeakForm, LCEDualWeakFormBCDirichlet,
                           LCEScaledDualWeakForm, LCEScaledDualWeakFormBCDirichlet, LERieszMapWeakForm,
                           LCERieszMapWeakForm, LCEScaledRieszMapWeakForm, LEDualWeakFormBCNeumann,
                           LCEDualWeakFormBCNeumann)
from .h1_multifield import (LCEPrimalWeakForm, LCEPrimalWeakFormBCDirichlet, LETaylorHoodWeakForm,
                            LETaylorHoodWeakFormBCDirichlet, LEPrimalStressConstraintWeakForm)
from .coupling_weak_form import CouplingWeakForm, interface_triples
