"""ccl_b200 -- B200-native Limber angular power spectra behind the pyccl API names.

Hot path: ``ccl_angular_cls_limber`` (reference src/ccl_cls.c:265-363) as hand-written sm_100a
CUDA kernels in ``libccl_b200.so``, reached through the C ABI of ``include/ccl_b200.h``.
"""
from . import _lib  # noqa: F401
from .lowlevel import (CCLB200Error, LibBatch, LibCosmology, LibF2D, LibF3D, LibTracer,  # noqa: F401
                       angular_cls_limber, make_params)

from .cells import angular_cl  # noqa: F401,E402
from .covariances import angular_cl_cov_cNG, angular_cl_cov_SSC  # noqa: F401,E402
from .tk3d import Tk3D  # noqa: F401,E402
from .correlations import CorrelationMethods, CorrelationTypes, correlation  # noqa: F401,E402
from .cosmology import (Cosmology, CosmologyCalculator, CosmologyVanillaLCDM, get_table_builder,  # noqa: F401,E402
                        set_table_builder)
from .errors import CCLError, CCLWarning, update_warning_verbosity  # noqa: F401,E402
from .parameters import DEFAULT_POWER_SPECTRUM, gsl_params, physical_constants, spline_params  # noqa: F401,E402
from .pipeline import PipelinedBatch, pin_stacked, slice_stacked  # noqa: F401,E402
from .pk2d import Pk2D, parse_pk2d  # noqa: F401,E402
from .tracers import (CMBLensingTracer, ISWTracer, NumberCountsTracer, NzTracer, Tracer,  # noqa: F401,E402
                      WeakLensingTracer, get_density_kernel, get_kappa_kernel, get_lensing_kernel)



def linear_matter_power(cosmo, k, a):
    return cosmo.linear_matter_power(k, a)


def nonlin_matter_power(cosmo, k, a):
    return cosmo.nonlin_matter_power(k, a)


def comoving_radial_distance(cosmo, a):
    return cosmo.comoving_radial_distance(a)


def scale_factor_of_chi(cosmo, chi):
    return cosmo.scale_factor_of_chi(chi)


def growth_factor(cosmo, a):
    return cosmo.growth_factor(a)


def growth_rate(cosmo, a):
    return cosmo.growth_rate(a)


def h_over_h0(cosmo, a):
    return cosmo.h_over_h0(a)


__version__ = "0.1.0"
