"""Global accuracy parameters and constants with the reference's names and snapshot semantics:
``spline_params`` / ``gsl_params`` are process-global, mutable, have ``.reload()``, and are
copied by value into each Cosmology at construction (pyccl/_core/parameters/parameters_base.py:
20-141; defaults src/ccl_core.c:67-141; snapshot src/ccl_core.c:247-248)."""
import copy
import math

__all__ = ("spline_params", "gsl_params", "physical_constants", "DEFAULT_POWER_SPECTRUM",
           "DefaultParams", "CCLParameters")

DEFAULT_POWER_SPECTRUM = "delta_matter:delta_matter"


class DefaultParams:
    T_CMB = 2.7255
    T_ncdm = 0.71611


class CCLParameters:
    _defaults = {}
    _frozen_cls = False

    def __init__(self):
        object.__setattr__(self, "_values", dict(self._defaults))
        object.__setattr__(self, "_frozen", self._frozen_cls)

    def __getattr__(self, name):
        vals = object.__getattribute__(self, "_values")
        if name in vals:
            return vals[name]
        raise AttributeError(name)

    def __setattr__(self, key, value):
        if self._frozen:
            raise AttributeError("Instances of %s are frozen." % type(self).__name__)
        if key not in self._values:
            raise KeyError("Parameter %s does not exist." % key)
        self._values[key] = value

    def __getitem__(self, key):
        return getattr(self, key)

    __setitem__ = __setattr__

    def __repr__(self):
        return repr(dict(self._values))

    def reload(self):
        object.__setattr__(self, "_values", dict(self._defaults))

    def freeze(self):
        object.__setattr__(self, "_frozen", True)

    def unfreeze(self):
        object.__setattr__(self, "_frozen", False)

    def snapshot(self):
        return copy.deepcopy(self._values)


class SplineParams(CCLParameters):
    """ccl_spline_params (src/ccl_core.c:94-141)."""
    _defaults = dict(
        A_SPLINE_NA=250, A_SPLINE_MIN=0.1, A_SPLINE_MINLOG_PK=0.01, A_SPLINE_MIN_PK=0.1,
        A_SPLINE_MINLOG_SM=0.01, A_SPLINE_MIN_SM=0.1, A_SPLINE_MAX=1.0, A_SPLINE_MINLOG=0.0001,
        A_SPLINE_NLOG=250, LOGM_SPLINE_DELTA=0.025, LOGM_SPLINE_NM=50, LOGM_SPLINE_MIN=6,
        LOGM_SPLINE_MAX=17, A_SPLINE_NA_SM=13, A_SPLINE_NLOG_SM=6, A_SPLINE_NA_PK=40,
        A_SPLINE_NLOG_PK=11, K_MAX_SPLINE=50, K_MAX=1E3, K_MIN=5E-5, DLOGK_INTEGRATION=0.025,
        DCHI_INTEGRATION=5., N_K=167, N_K_3DCOR=100000, ELL_MIN_CORR=0.01, ELL_MAX_CORR=60000,
        N_ELL_CORR=5000)

    def __setattr__(self, key, value):
        if key == "A_SPLINE_MAX" and value != 1.0:
            raise ValueError("A_SPLINE_MAX is fixed to 1.")
        super().__setattr__(key, value)

    __setitem__ = __setattr__


class GSLParams(CCLParameters):
    """ccl_gsl_params (src/ccl_core.c:67-82).

    The Gauss-Kronrod entries hold GSL's rule key like the reference (GSL_INTEG_GAUSS41 = 4,
    src/ccl_core.c:40), not a node count.  Only GK41 exists on the device: a 'qag_quad' angular_cl
    call on a cosmology created with another INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS raises CCLError
    (CCL_ERROR_NOT_IMPLEMENTED) instead of silently using GK41."""
    _defaults = dict(
        N_ITERATION=1000, INTEGRATION_GAUSS_KRONROD_POINTS=4, INTEGRATION_EPSREL=1E-4,
        INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS=4, INTEGRATION_LIMBER_EPSREL=1E-4,
        INTEGRATION_DISTANCE_EPSREL=1E-6, INTEGRATION_SIGMAR_EPSREL=1E-5,
        INTEGRATION_KNL_EPSREL=1E-5, ROOT_EPSREL=1E-4, ROOT_N_ITERATION=1000,
        ODE_GROWTH_EPSREL=1E-6, EPS_SCALEFAC_GROWTH=1E-6, NZ_NORM_SPLINE_INTEGRATION=True,
        LENSING_KERNEL_SPLINE_INTEGRATION=True)


class PhysicalConstants(CCLParameters):
    """ccl_constants (src/ccl_core.c:144-219)."""
    _frozen_cls = True
    _defaults = dict(
        YEAR=365.256363004 * 86400., CLIGHT_HMPC=2997.92458, GNEWT=6.67430e-11,
        SOLAR_MASS=1.988409871E+30, MPC_TO_METER=3.085677581491367399198952281E+22,
        RHO_CRITICAL=((3 * 100 * 100) / (8 * math.pi * 6.67430e-11)) *
                     (1000 * 1000 * 3.085677581491367399198952281E+22 / 1.988409871E+30),
        KBOLTZ=1.380649E-23, STBOLTZ=5.670374419E-8, HPLANCK=6.62607015E-34, CLIGHT=299792458.0,
        EV_IN_J=1.602176634E-19, DELTAM12_sq=7.62E-5, DELTAM13_sq_pos=2.55E-3,
        DELTAM13_sq_neg=-2.43E-3)


spline_params = SplineParams()
gsl_params = GSLParams()
physical_constants = PhysicalConstants()
