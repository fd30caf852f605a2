"""``Cosmology`` with the pyccl constructor, item access and lazy ``compute_*`` methods
(pyccl/cosmology.py:134-751), restricted to what the Limber path reads.

The object owns (i) the derived parameter set (pyccl/cosmology.py:398-496), (ii) host tables of
the background (src/ccl_background.c) and of the analytic matter power spectra
(src/ccl_power.c, ccl_bbks.c, ccl_eh.c, ccl_halofit.c) built by ``background.py`` /
``power.py``, and (iii) the device-side ``ccl_b200_cosmology`` handle the C ABI needs, created on
first use.  Boltzmann codes, emulators, massive neutrinos and modified gravity are out of scope
(SURVEY.md section 2): asking for them raises immediately instead of silently approximating.
"""
import numpy as np

from . import lowlevel
from ._hostspline import Akima
from .background import Background, h_over_h0, omega_x
from .errors import CCLError
from .parameters import DEFAULT_POWER_SPECTRUM, DefaultParams, gsl_params, physical_constants, spline_params

__all__ = ("Cosmology", "CosmologyVanillaLCDM", "CosmologyCalculator", "set_table_builder", "get_table_builder")

# Who builds the tables of a Cosmology (background, growth, analytic P(k), halofit, tracer kernels):
#   "device": the ccl_b200_build_* kernels (csrc/builders.cu), one-cosmology batches -- the product path; it fails
#             loudly without a CUDA device, like every other compute entry;
#   "host":   the numpy twins (background.py, power.py, tracers.py).  They are the test oracles of the device
#             builders (pinned on the reference's benchmark files, tests/test_upstream_tables_cpu.py) and what
#             the CPU-only test session uses to prepare inputs for the oracle; tests select them explicitly.
_TABLE_BUILDER = "device"


def set_table_builder(which):
    global _TABLE_BUILDER
    if which not in ("device", "host"):
        raise ValueError("table builder must be 'device' or 'host'")
    _TABLE_BUILDER = which


def get_table_builder():
    return _TABLE_BUILDER


_ANALYTIC_TRANSFER = ("bbks", "eisenstein_hu", "eisenstein_hu_nowiggles")
_TRANSFER_TYPES = _ANALYTIC_TRANSFER + ("boltzmann_camb", "boltzmann_class", "boltzmann_isitgr", "calculator")
_MPS_TYPES = ("linear", "halofit", "camb", "calculator")


class Cosmology:
    """Same keyword-only signature and defaults as ``pyccl.Cosmology`` (pyccl/cosmology.py:235-245)."""

    def __init__(self, *, Omega_c=None, Omega_b=None, h=None, n_s=None, sigma8=None, A_s=None,
                 Omega_k=0., Omega_g=None, Neff=None, m_nu=0., mass_split='normal', w0=-1., wa=0.,
                 T_CMB=DefaultParams.T_CMB, T_ncdm=DefaultParams.T_ncdm,
                 transfer_function='boltzmann_camb', matter_power_spectrum='halofit',
                 baryonic_effects=None, mg_parametrization=None, extra_parameters=None):
        if Neff is None:
            Neff = 3.044
        if transfer_function not in _TRANSFER_TYPES:
            raise ValueError("Unknown transfer_function %r" % (transfer_function,))
        if matter_power_spectrum not in _MPS_TYPES:
            raise ValueError("Unknown matter_power_spectrum %r" % (matter_power_spectrum,))
        if baryonic_effects is not None or mg_parametrization is not None:
            raise NotImplementedError("baryonic effects / modified gravity are outside the Limber hot path")
        self._config_init_kwargs = dict(transfer_function=transfer_function,
                                        matter_power_spectrum=matter_power_spectrum,
                                        extra_parameters=extra_parameters)
        self.transfer_function_type = transfer_function
        self.matter_power_spectrum_type = matter_power_spectrum
        self._params_init_kwargs = dict(Omega_c=Omega_c, Omega_b=Omega_b, h=h, n_s=n_s, sigma8=sigma8, A_s=A_s,
                                        Omega_k=Omega_k, Omega_g=Omega_g, Neff=Neff, m_nu=m_nu,
                                        mass_split=mass_split, w0=w0, wa=wa, T_CMB=T_CMB, T_ncdm=T_ncdm)
        self._params = self._build_parameters(**self._params_init_kwargs)
        # accuracy parameters are snapshotted by value at creation (src/ccl_core.c:247-248)
        self._spline_params = spline_params.snapshot()
        self._gsl_params = gsl_params.snapshot()
        self._bg = Background(self._params, self._spline_params, self._gsl_params)
        self._pk_lin = {}
        self._pk_nl = {}
        self._has_distances = self._has_growth = False
        self._dev = None
        self._splines = {}
        self._table_builder = _TABLE_BUILDER   # snapshotted like the accuracy parameters
        self._bgd = None                        # device-built background tables (device_builders.build_background)

    # ---- parameters (pyccl/cosmology.py:398-496, m_nu = 0 branch) ----
    @staticmethod
    def _build_parameters(Omega_c, Omega_b, h, n_s, sigma8, A_s, Omega_k, Omega_g, Neff, m_nu, mass_split, w0,
                          wa, T_CMB, T_ncdm):
        A_s = np.nan if A_s is None else A_s
        sigma8 = np.nan if sigma8 is None else sigma8
        Omega_g_in = np.nan if Omega_g is None else Omega_g
        if Omega_k < -1.0135:
            raise ValueError("Omega_k must be more than -1.0135.")
        if [np.isnan(A_s), np.isnan(sigma8)].count(True) != 1:
            raise ValueError("Set either A_s or sigma8 but not both.")
        for name, value in (("Omega_c", Omega_c), ("Omega_b", Omega_b), ("h", h), ("n_s", n_s)):
            if value is None:
                raise ValueError("Must set parameter %s." % name)
        if np.any(np.atleast_1d(m_nu) != 0):
            raise NotImplementedError("massive neutrinos are outside the Limber hot path (SURVEY.md 2a)")
        c = physical_constants
        k_sign = -np.sign(Omega_k) if np.abs(Omega_k) > 1e-6 else 0
        sqrtk = np.sqrt(np.abs(Omega_k)) * h / c.CLIGHT_HMPC
        rho_g = 4 * c.STBOLTZ / c.CLIGHT ** 3 * T_CMB ** 4
        rho_crit = c.RHO_CRITICAL * c.SOLAR_MASS / c.MPC_TO_METER ** 3 * h ** 2
        g = (4 / 11) ** (1 / 3)
        T_nu = g * T_CMB
        N_nu_rel = Neff
        rho_nu_rel = N_nu_rel * (7 / 8) * 4 * c.STBOLTZ / c.CLIGHT ** 3 * T_nu ** 4
        Omega_nu_rel = rho_nu_rel / rho_crit
        Omega_m = Omega_b + Omega_c
        Omega_l = 1 - Omega_m - rho_g / rho_crit - Omega_nu_rel - Omega_k
        if np.isnan(Omega_g_in):
            Omega_g_val = rho_g / rho_crit
        else:
            Omega_g_val = Omega_g_in
            Omega_l += rho_g / rho_crit - Omega_g_in
        return dict(m_nu=0.0, sum_nu_masses=0.0, N_nu_mass=0, N_nu_rel=N_nu_rel, Neff=Neff, Omega_nu_mass=0.0,
                    Omega_nu_rel=Omega_nu_rel, Omega_m=Omega_m, Omega_c=Omega_c, Omega_b=Omega_b, Omega_k=Omega_k,
                    sqrtk=sqrtk, k_sign=int(k_sign), T_CMB=T_CMB, T_ncdm=T_ncdm, Omega_g=Omega_g_val, w0=w0,
                    wa=wa, Omega_l=Omega_l, h=h, H0=h * 100, A_s=A_s, sigma8=sigma8, n_s=n_s, mu_0=0.0,
                    sigma_0=0.0, c1_mg=1.0, c2_mg=1.0, lambda_mg=0.0)

    def __getitem__(self, key):
        if key == 'extra_parameters':
            return self._config_init_kwargs['extra_parameters']
        if key not in self._params:
            raise KeyError("Parameter %s not recognized." % key)
        return self._params[key]

    def __repr__(self):
        kw = ", ".join("%s=%r" % kv for kv in {**self._params_init_kwargs, **self._config_init_kwargs}.items())
        return "ccl_b200.Cosmology(%s)" % kw

    # pickling by constructor kwargs, like the reference (pyccl/cosmology.py:532-548)
    def __getstate__(self):
        return {**self._params_init_kwargs, **self._config_init_kwargs}

    def __setstate__(self, state):
        self.__init__(**state)

    # ---- lazy table builders ----
    @property
    def has_distances(self):
        return self._has_distances

    @property
    def has_growth(self):
        return self._has_growth

    @property
    def has_linear_power(self):
        return DEFAULT_POWER_SPECTRUM in self._pk_lin

    @property
    def has_nonlin_power(self):
        return DEFAULT_POWER_SPECTRUM in self._pk_nl

    def compute_distances(self):
        """ccl_cosmology_compute_distances (src/ccl_background.c:417-589)."""
        if self._has_distances:
            return
        if self._table_builder == "device":
            self._device_background()
        else:
            self._bg.compute_distances()
        self._splines["chi"] = Akima(self._bg.a, self._bg.chi)
        self._splines["E"] = Akima(self._bg.a, self._bg.E)
        self._splines["achi"] = Akima(self._bg.achi_chi, self._bg.achi_a)
        self._has_distances = True

    def compute_growth(self):
        """ccl_cosmology_compute_growth (src/ccl_background.c:763-1002)."""
        if self._has_growth:
            return
        if self._table_builder == "device":
            self._device_background()
        else:
            self._bg.compute_growth()
        self._splines["D"] = Akima(self._bg.a, self._bg.growth)
        self._splines["f"] = Akima(self._bg.a, self._bg.fgrowth)
        self._has_growth = True

    def _device_background(self):
        """E(a), chi(a), a(chi), D(a), f(a) in one launch sequence of the device builders
        (ccl_b200_build_background; src/ccl_background.c:417-589,763-1002)."""
        if self._bgd is not None:
            return
        from . import device_builders as db
        bgd = db.build_background([self], self._spline_params, self._gsl_params)
        n = int(bgd["n_achi"][0])
        b = self._bg
        b.E, b.chi = bgd["E"][0], bgd["chi"][0]
        b.achi_chi, b.achi_a = np.ascontiguousarray(bgd["achi_chi"][0, :n]), np.ascontiguousarray(bgd["achi_a"][0, :n])
        b.growth, b.fgrowth, b.growth0 = bgd["growth"][0], bgd["fgrowth"][0], float(bgd["growth0"][0])
        self._bgd = bgd

    def compute_linear_power(self):
        """pyccl/cosmology.py:558-613 for the analytic transfer functions."""
        if self.has_linear_power:
            return
        from .pk2d import Pk2D
        trf = self.transfer_function_type
        if trf not in _ANALYTIC_TRANSFER:
            raise CCLError("transfer_function=%r needs an external Boltzmann code, which this build does not "
                           "wrap; use 'bbks', 'eisenstein_hu' or 'eisenstein_hu_nowiggles', or a "
                           "CosmologyCalculator / Pk2D with precomputed arrays" % trf)
        if self._table_builder == "device":
            from . import device_builders as db
            if np.isnan(self["sigma8"]):
                raise CCLError("Error CCL_ERROR_INCONSISTENT: sigma8 not set, required for analytic power spectra")
            self.compute_growth()
            lin = db.build_linear_power([self], trf, self._bgd, self._spline_params)
            self._pk_lin[DEFAULT_POWER_SPECTRUM] = Pk2D(a_arr=lin["a"], lk_arr=lin["lk"], pk_arr=lin["logp"][0], is_logp=True,
                                                        extrap_order_lok=1, extrap_order_hik=2)
        else:
            self._pk_lin[DEFAULT_POWER_SPECTRUM] = Pk2D.from_model(self, trf)

    def compute_nonlin_power(self):
        """pyccl/cosmology.py:622-647."""
        if self.has_nonlin_power:
            return
        self.compute_linear_power()
        mps = self.matter_power_spectrum_type
        pkl = self._pk_lin[DEFAULT_POWER_SPECTRUM]
        if mps == 'linear':
            self._pk_nl[DEFAULT_POWER_SPECTRUM] = pkl
        elif mps == 'halofit':
            self._pk_nl[DEFAULT_POWER_SPECTRUM] = pkl.apply_halofit(self)
        else:
            raise CCLError("matter_power_spectrum=%r is not available in this build" % mps)

    def get_linear_power(self, name=DEFAULT_POWER_SPECTRUM):
        if name == DEFAULT_POWER_SPECTRUM:
            self.compute_linear_power()
        if name not in self._pk_lin:
            raise KeyError("Power spectrum %s does not exist." % name)
        return self._pk_lin[name]

    def get_nonlin_power(self, name=DEFAULT_POWER_SPECTRUM):
        if name == DEFAULT_POWER_SPECTRUM:
            self.compute_nonlin_power()
        if name not in self._pk_nl:
            raise KeyError("Power spectrum %s does not exist." % name)
        return self._pk_nl[name]

    def parse_pk2d(self, p_of_k_a=DEFAULT_POWER_SPECTRUM, *, is_linear=False):
        """pyccl/pk2d.py:509-541."""
        from .pk2d import Pk2D
        if isinstance(p_of_k_a, Pk2D):
            return p_of_k_a
        if isinstance(p_of_k_a, str):
            return self.get_linear_power(p_of_k_a) if is_linear else self.get_nonlin_power(p_of_k_a)
        raise ValueError("p_of_k_a must be a pyccl.Pk2D object or a string.")

    parse_pk = parse_pk2d

    # ---- background accessors used by the tracers (pyccl/background.py) ----
    def h_over_h0(self, a):
        return h_over_h0(a, self._params) + 0.0 * np.asarray(a, dtype=float)

    def omega_x(self, a, species):
        return omega_x(a, self._params, species)

    def comoving_radial_distance(self, a):
        """ccl_comoving_radial_distance: Akima spline of chi(a), 0 at a = 1."""
        self.compute_distances()
        a_arr = np.asarray(a, dtype=float)
        if np.any(a_arr > 1.0) or np.any(a_arr < self._bg.a[0]):
            raise CCLError("Error CCL_ERROR_COMPUTECHI: scale factor outside the interpolation range [%g, 1]"
                           % self._bg.a[0])
        out = np.where(a_arr == 1.0, 0.0, self._splines["chi"](a_arr))
        return out if out.ndim else float(out)

    def scale_factor_of_chi(self, chi):
        """ccl_scale_factor_of_chi (src/ccl_background.c:1251-1277) on the host table."""
        self.compute_distances()
        c = np.asarray(chi, dtype=float)
        if np.any(c < 0) or np.any(c > self._bg.achi_chi[-1]):
            raise CCLError("Error CCL_ERROR_COMPUTECHI: chi outside the a(chi) table")
        out = np.where(c < 1e-8, 1.0, self._splines["achi"](np.maximum(c, 0.0)))
        return out if out.ndim else float(out)

    def growth_factor(self, a):
        self.compute_growth()
        a_arr = np.asarray(a, dtype=float)
        out = np.where(a_arr == 1.0, 1.0, self._splines["D"](a_arr))
        return out if out.ndim else float(out)

    def growth_factor_unnorm(self, a):
        return self.growth_factor(a) * self._bg.growth0

    def growth_rate(self, a):
        self.compute_growth()
        out = self._splines["f"](np.asarray(a, dtype=float))
        return out if out.ndim else float(out)

    def linear_matter_power(self, k, a):
        """pyccl.linear_matter_power (pyccl/power.py): evaluated on the device."""
        return self.get_linear_power()(k, a, self)

    def nonlin_matter_power(self, k, a):
        """pyccl.nonlin_matter_power (pyccl/power.py): evaluated on the device."""
        return self.get_nonlin_power()(k, a, self)

    def sigma8(self, p_of_k_a=DEFAULT_POWER_SPECTRUM):
        pk = self.parse_pk2d(p_of_k_a, is_linear=True)
        from .power import sigmaR_from_table
        a, lk, tab = pk.get_spline_arrays_raw()
        if not pk.is_logp:
            tab = np.log(tab)
        return float(sigmaR_from_table(lk, tab[-1], 8.0 / self["h"], self._spline_params))

    # ---- device side ----
    def _limber_params(self):
        sp, gp = self._spline_params, self._gsl_params
        return lowlevel.make_params(K_MAX=sp["K_MAX"], K_MIN=sp["K_MIN"], DLOGK_INTEGRATION=sp["DLOGK_INTEGRATION"],
                                    INTEGRATION_LIMBER_EPSREL=gp["INTEGRATION_LIMBER_EPSREL"],
                                    N_ITERATION=gp["N_ITERATION"], A_SPLINE_MIN=sp["A_SPLINE_MIN"],
                                    INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS=gp[
                                        "INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS"],
                                    DCHI_INTEGRATION=sp["DCHI_INTEGRATION"], ELL_MIN_CORR=sp["ELL_MIN_CORR"],
                                    ELL_MAX_CORR=sp["ELL_MAX_CORR"], N_ELL_CORR=sp["N_ELL_CORR"],
                                    INTEGRATION_GAUSS_KRONROD_POINTS=gp["INTEGRATION_GAUSS_KRONROD_POINTS"],
                                    INTEGRATION_EPSREL=gp["INTEGRATION_EPSREL"])

    def chi_max_nokernel(self):
        """chi(A_SPLINE_MIN): chi_max of kernel-less tracers (src/ccl_tracers.c:632)."""
        return float(self.comoving_radial_distance(self._spline_params["A_SPLINE_MIN"]))

    def _device(self, need_growth=False):
        """The ccl_b200_cosmology handle (a(chi) + growth tables in HBM); created on first use."""
        self.compute_distances()
        if self._dev is None:
            dev = lowlevel.LibCosmology(self._limber_params())
            dev.set_achi(self._bg.achi_chi, self._bg.achi_a)
            dev.set_chi_max_nokernel(self.chi_max_nokernel())
            self._dev = dev
            self._dev_has_growth = False
        if need_growth and not self._dev_has_growth:
            self.compute_growth()
            self._dev.set_growth(self._bg.a, self._bg.growth)
            self._dev_has_growth = True
        return self._dev

    def angular_cl(self, *args, **kwargs):
        from .cells import angular_cl
        return angular_cl(self, *args, **kwargs)


def CosmologyVanillaLCDM(**kwargs):
    """pyccl/cosmology.py:730-751."""
    p = {'Omega_c': 0.25, 'Omega_b': 0.05, 'h': 0.67, 'n_s': 0.96, 'sigma8': 0.81, 'A_s': None}
    if any(k in kwargs for k in p):
        raise ValueError("You cannot change the LCDM parameters: %s " % list(p.keys()))
    return Cosmology(**{**p, **kwargs})


class CosmologyCalculator(Cosmology):
    """Calculator mode (pyccl/cosmology.py:868-1001): background, growth and P(k) tables are
    injected instead of computed (ccl_cosmology_distances_from_input / _growth_from_input,
    src/ccl_background.c:595-756)."""

    def __init__(self, *, Omega_c=None, Omega_b=None, h=None, n_s=None, sigma8=None, A_s=None, Omega_k=0.,
                 Omega_g=None, Neff=None, m_nu=0., mass_split='normal', w0=-1., wa=0.,
                 T_CMB=DefaultParams.T_CMB, T_ncdm=DefaultParams.T_ncdm, background=None, growth=None,
                 pk_linear=None, pk_nonlin=None, nonlinear_model=None):
        super().__init__(Omega_c=Omega_c, Omega_b=Omega_b, h=h, n_s=n_s, sigma8=sigma8, A_s=A_s, Omega_k=Omega_k,
                         Omega_g=Omega_g, Neff=Neff, m_nu=m_nu, mass_split=mass_split, w0=w0, wa=wa,
                         T_CMB=T_CMB, T_ncdm=T_ncdm, transfer_function='calculator',
                         matter_power_spectrum='calculator')
        # injected tables: whatever is still computed (e.g. growth when only distances are given) must stay consistent
        # with them, so the calculator keeps the host twins and never overwrites its background from the device
        self._table_builder = "host"
        self._input_arrays = dict(background=background, growth=growth, pk_linear=pk_linear, pk_nonlin=pk_nonlin,
                                  nonlinear_model=nonlinear_model)
        if background is not None:
            self._init_background(background)
        if growth is not None:
            self._init_growth(growth)
        if pk_linear is not None:
            self._init_pk(pk_linear, self._pk_lin)
        if pk_nonlin is not None:
            self._init_pk(pk_nonlin, self._pk_nl)
        if nonlinear_model is not None:
            if nonlinear_model != 'halofit' or pk_linear is None:
                raise ValueError("nonlinear_model must be 'halofit' (with pk_linear) or None in this build")
            for name, pk in self._pk_lin.items():
                if name not in self._pk_nl:
                    self._pk_nl[name] = pk.apply_halofit(self)

    def __getstate__(self):
        raise TypeError("CosmologyCalculator holds user tables and is not re-creatable from kwargs")

    @staticmethod
    def _check_a(a):
        a = np.asarray(a, dtype=float)
        if a[-1] != 1.0 or np.any(np.diff(a) <= 0):
            raise ValueError("Input scale factor array must be monotonically increasing and end at a = 1")
        return a

    def _init_background(self, bg):
        if not isinstance(bg, dict) or not {'a', 'chi', 'h_over_h0'} <= set(bg):
            raise ValueError("`background` must be a dictionary with keys 'a', 'chi' and 'h_over_h0'")
        a = self._check_a(bg['a'])
        chi = np.asarray(bg['chi'], dtype=float)
        E = np.asarray(bg['h_over_h0'], dtype=float)
        if chi.shape != a.shape or E.shape != a.shape:
            raise ValueError("Input arrays for the background must have the same shape")
        if np.any(np.diff(chi) >= 0):
            raise ValueError("Comoving distance must decrease with the scale factor")
        b = self._bg
        b.a, b.chi, b.E = a, chi, E
        # ccl_cosmology_distances_from_input: a(chi) is the spline of the reversed input arrays
        b.achi_chi, b.achi_a = chi[::-1].copy(), a[::-1].copy()
        self._spline_params = dict(self._spline_params, A_SPLINE_MIN=max(self._spline_params["A_SPLINE_MIN"], a[0]))
        self._splines["chi"] = Akima(a, chi)
        self._splines["E"] = Akima(a, E)
        self._splines["achi"] = Akima(b.achi_chi, b.achi_a)
        self._has_distances = True

    def h_over_h0(self, a):
        if "E" in self._splines and self._input_arrays.get("background") is not None:
            out = self._splines["E"](np.asarray(a, dtype=float))
            return out if out.ndim else float(out)
        return super().h_over_h0(a)

    def _init_growth(self, gr):
        if not isinstance(gr, dict) or not {'a', 'growth_factor', 'growth_rate'} <= set(gr):
            raise ValueError("`growth` must be a dictionary with keys 'a', 'growth_factor' and 'growth_rate'")
        a = self._check_a(gr['a'])
        D = np.asarray(gr['growth_factor'], dtype=float)
        f = np.asarray(gr['growth_rate'], dtype=float)
        self._growth_a = a
        self._bg.growth0 = D[-1]
        self._bg.growth, self._bg.fgrowth = D / D[-1], f
        self._splines["D"] = Akima(a, D / D[-1])
        self._splines["f"] = Akima(a, f)
        self._has_growth = True

    def _device(self, need_growth=False):
        dev = super()._device(False)
        if need_growth and not self._dev_has_growth:
            a = getattr(self, "_growth_a", self._bg.a)
            self.compute_growth()
            dev.set_growth(a, self._bg.growth)
            self._dev_has_growth = True
        return dev

    def _device_background(self):
        """E(a), chi(a), a(chi), D(a), f(a) in one launch sequence of the device builders
        (ccl_b200_build_background; src/ccl_background.c:417-589,763-1002)."""
        if self._bgd is not None:
            return
        from . import device_builders as db
        bgd = db.build_background([self], self._spline_params, self._gsl_params)
        n = int(bgd["n_achi"][0])
        b = self._bg
        b.E, b.chi = bgd["E"][0], bgd["chi"][0]
        b.achi_chi, b.achi_a = np.ascontiguousarray(bgd["achi_chi"][0, :n]), np.ascontiguousarray(bgd["achi_a"][0, :n])
        b.growth, b.fgrowth, b.growth0 = bgd["growth"][0], bgd["fgrowth"][0], float(bgd["growth0"][0])
        self._bgd = bgd

    def compute_linear_power(self):
        if not self.has_linear_power:
            raise CCLError("Error CCL_ERROR_LINEAR_POWER_INIT: the calculator was given no pk_linear")

    def compute_nonlin_power(self):
        if not self.has_nonlin_power:
            raise CCLError("Error CCL_ERROR_NONLIN_POWER_INIT: the calculator was given no pk_nonlin")

    def _init_pk(self, spec, store):
        from .pk2d import Pk2D
        if not isinstance(spec, dict) or not {'a', 'k', DEFAULT_POWER_SPECTRUM} <= set(spec):
            raise ValueError("power spectrum dictionaries need keys 'a', 'k' and 'delta_matter:delta_matter'")
        a = np.asarray(spec['a'], dtype=float)
        lk = np.log(np.asarray(spec['k'], dtype=float))
        for name, pk in spec.items():
            if name in ('a', 'k'):
                continue
            pk = np.asarray(pk, dtype=float)
            if pk.shape != (a.size, lk.size):
                raise ValueError("The power spectrum %s has shape %s but shape %s was expected."
                                 % (name, pk.shape, (a.size, lk.size)))
            if np.any(pk <= 0):
                store[name] = Pk2D(a_arr=a, lk_arr=lk, pk_arr=pk, is_logp=False)
            else:
                store[name] = Pk2D(a_arr=a, lk_arr=lk, pk_arr=np.log(pk), is_logp=True)
