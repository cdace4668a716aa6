"""Python mirror of the reference host API (`namespace kinetix`, reference benchmark/src/kinetix.hpp:15-107)
on top of the C ABI in include/kinetix_b200.h (ctypes; no torch types cross the boundary).

Same names, argument meaning and call order as the reference so that tests read like its benchmark
driver (benchmark/src/bk.cpp:574-760):

    import kinetix_b200.host as kinetix
    kinetix.init("gri30.yaml", device_id=0)
    kinetix.build(p_ref, T_ref, Y_ref, transport=True)
    kinetix.productionRates(n_states, offsetT, offset, p / p_ref, state, rates)
    kinetix.mixtureAvgTransportProps(n_states, offsetT, offset, p / p_ref, state, viscosity, conductivity, rhoD)

Device buffers may be torch CUDA tensors (float64, their data_ptr() is passed) or raw integer device
addresses.  Errors: the reference asserts / aborts; here every non-zero status raises KinetixError with
the library's message.  There is no CPU fallback: without the compiled CUDA library this module raises.
"""
import ctypes
import os
import sys

from . import jit

KX_DTYPE_F64 = 0
KX_DTYPE_F32 = 1


class KinetixError(RuntimeError):
    pass


class _Options(ctypes.Structure):
    _fields_ = [('device_id', ctypes.c_int), ('block_size', ctypes.c_int), ('single_precision', ctypes.c_int),
                ('unroll_loops', ctypes.c_int), ('loop_gibbsexp', ctypes.c_int), ('group_rxn_unroll', ctypes.c_int),
                ('group_vis', ctypes.c_int), ('nonsym_dij', ctypes.c_int), ('fit_rcp_diff_coeffs', ctypes.c_int),
                ('verbose', ctypes.c_int), ('cache_dir', ctypes.c_char_p), ('tool', ctypes.c_char_p)]


_lib = None
_i64, _dbl, _vp, _int = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int


def library():
    """Load libkinetix_b200.so (built in-tree by __graft_entry__.build() / jit.build_host_library())."""
    global _lib
    if _lib is not None:
        return _lib
    path = jit.HOST_LIB
    if not os.path.exists(path):
        raise KinetixError(f'{path} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           f'(kinetix_b200 has no CPU fallback)')
    lib = ctypes.CDLL(path)
    lib.kx_init.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Options)]
    lib.kx_build.argtypes = [_dbl, _dbl, ctypes.POINTER(_dbl), _int]
    lib.kx_production_rates.argtypes = [_i64, _i64, _i64, _dbl, _vp, _vp, _int, _vp]
    lib.kx_mixture_avg_transport_props.argtypes = [_i64, _i64, _i64, _dbl, _vp, _vp, _vp, _vp, _int, _vp]
    lib.kx_thermodynamic_props.argtypes = [_i64, _i64, _i64, _dbl, _vp, _vp, _vp, _vp, _int, _vp]
    lib.kx_production_rates_pfield.argtypes = [_i64, _i64, _i64, _vp, _vp, _vp, _int, _vp]
    lib.kx_thermodynamic_props_pfield.argtypes = [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _vp]
    lib.kx_production_rates_host.argtypes = [_i64, _i64, _i64, _dbl, _vp, _vp]
    lib.kx_mixture_avg_transport_props_host.argtypes = [_i64, _i64, _i64, _dbl, _vp, _vp, _vp, _vp]
    lib.kx_rates_and_transport_host.argtypes = [_i64, _i64, _i64, _dbl, _vp, _vp, _vp, _vp, _vp]
    lib.kx_species_name.argtypes = [_int]
    lib.kx_species_name.restype = ctypes.c_char_p
    lib.kx_species_index.argtypes = [ctypes.c_char_p]
    for f in ('kx_molecular_weights', 'kx_molar_masses', 'kx_ref_mass_fractions'):
        getattr(lib, f).argtypes = [ctypes.POINTER(_dbl)]
    for f in ('kx_ref_pressure', 'kx_ref_temperature', 'kx_ref_mean_molecular_weight'):
        getattr(lib, f).restype = _dbl
    lib.kx_last_error.restype = ctypes.c_char_p
    lib.kx_module_path.restype = ctypes.c_char_p
    lib.kx_select_device.argtypes = [_int]
    lib.kx_fnv1a64.restype = ctypes.c_uint64
    lib.kx_fnv1a64.argtypes = [ctypes.c_uint64, ctypes.c_char_p, ctypes.c_size_t]
    _lib = lib
    return lib


def _check(status, what):
    if status != 0:
        raise KinetixError(f'{what}: {library().kx_last_error().decode()} (status {status})')


def _ptr(buf):
    """device (or host) address of a torch tensor / numpy array / int."""
    if buf is None:
        return None
    if hasattr(buf, 'data_ptr'):
        return buf.data_ptr()
    if hasattr(buf, 'ctypes'):
        return buf.ctypes.data
    return int(buf)


def _stream(stream):
    if stream is None:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.cuda.current_stream().cuda_stream
        except ImportError:
            pass
        return None
    return getattr(stream, 'cuda_stream', stream)


# ---- kinetix:: API -----------------------------------------------------------------------------------
def selectDevice(device_id):
    """extension (kx_select_device): make the context initialised on `device_id` the calling thread's current one;
    the reference drives one device per process (kinetix.cpp:21-65)"""
    _check(library().kx_select_device(int(device_id)), 'kinetix.selectDevice')


def currentDevice():
    return int(library().kx_current_device())


def isInitialized():
    return bool(library().kx_is_initialized())


def init(yamlPath, device_id=0, tool='KinetiX', blockSize=0, single_precision=False, unroll_loops=False,
         loop_gibbsexp=False, group_rxnUnroll=False, group_vis=False, nonsymDij=False, fit_rcpDiffCoeffs=False,
         align_width=0, target='sm_100a', useFP64Transport=False, verbose=False, cache_dir=None):
    """kinetix::init (kinetix.hpp:17-35).  occa::device -> device_id; MPI_Comm dropped; align_width,
    target and useFP64Transport are accepted and ignored (the reference ignores the last one too)."""
    o = _Options(device_id=device_id, block_size=blockSize, single_precision=int(single_precision),
                 unroll_loops=int(unroll_loops), loop_gibbsexp=int(loop_gibbsexp),
                 group_rxn_unroll=int(group_rxnUnroll), group_vis=int(group_vis), nonsym_dij=int(nonsymDij),
                 fit_rcp_diff_coeffs=int(fit_rcpDiffCoeffs), verbose=int(verbose),
                 cache_dir=cache_dir.encode() if cache_dir else None, tool=tool.encode() if tool else None)
    _check(library().kx_init(os.fspath(yamlPath).encode(), ctypes.byref(o)), 'kinetix.init')


_BUILD_MODULE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(_Options), ctypes.c_char_p, ctypes.c_void_p)
_builder_ref = None          # keeps the ctypes trampoline alive while it is installed


def setBuildKernel(builder):
    """Install a host-supplied module builder -- the counterpart of init's optional `buildKernel` argument
    (kinetix.hpp:11-13, kinetix.cpp:499-502,542).  `builder(yaml_path: str, options: dict, output_dir: str) -> int`
    is called by init / prepare instead of the built-in generator when the module is not cached; it must leave
    `output_dir/libkx_mech.so` and return 0.  None restores the default."""
    global _builder_ref
    lib = library()
    lib.kx_set_module_builder.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    if builder is None:
        lib.kx_set_module_builder(None, None)
        _builder_ref = None
        return

    def trampoline(yaml_path, opt, output_dir, _user):
        try:
            o = opt.contents
            options = {name: getattr(o, name) for name, _ in _Options._fields_ if name not in ('cache_dir', 'tool')}
            return int(builder(yaml_path.decode(), options, output_dir.decode()) or 0)
        except Exception as e:          # never let an exception cross the C boundary
            sys.stderr.write(f'[kinetix_b200] module builder raised: {e!r}\n')
            return -1
    ref = _BUILD_MODULE_FN(trampoline)
    lib.kx_set_module_builder(ctypes.cast(ref, ctypes.c_void_p), None)
    _builder_ref = ref


def prepare(yamlPath, single_precision=False, fit_rcpDiffCoeffs=False, blockSize=0, cache_dir=None, verbose=False):
    """kx_prepare: generate + compile the module into the cache without touching CUDA (what rank 0 does before
    the other ranks start, kinetix.cpp:290-296,655-699)."""
    o = _Options(block_size=blockSize, single_precision=int(single_precision),
                 fit_rcp_diff_coeffs=int(fit_rcpDiffCoeffs), verbose=int(verbose),
                 cache_dir=cache_dir.encode() if cache_dir else None)
    lib = library()
    lib.kx_prepare.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Options)]
    _check(lib.kx_prepare(os.fspath(yamlPath).encode(), ctypes.byref(o)), 'kinetix.prepare')


def build(refPressure, refTemperature, refMassFractions, transport=True):
    """kinetix::build (kinetix.hpp:58-63)."""
    n = nSpecies()
    if len(refMassFractions) != n:
        raise KinetixError(f'kinetix.build: refMassFractions has {len(refMassFractions)} entries, expected {n}')
    arr = (_dbl * n)(*[float(v) for v in refMassFractions])
    _check(library().kx_build(refPressure, refTemperature, arr, int(transport)), 'kinetix.build')


def productionRates(n_states, offsetT, offset, pressure, o_state, o_rates, stream=None, dtype=KX_DTYPE_F64):
    """kinetix::productionRates (kinetix.hpp:65-72); `pressure` is p / p_ref."""
    _check(library().kx_production_rates(n_states, offsetT, offset, pressure, _ptr(o_state), _ptr(o_rates), dtype,
                                         _stream(stream)), 'kinetix.productionRates')


def mixtureAvgTransportProps(nStates, offsetT, offset, pressure, o_state, o_viscosity, o_thermalConductivity,
                             o_densityDiffCoeffs, stream=None, dtype=KX_DTYPE_F64):
    """kinetix::mixtureAvgTransportProps (kinetix.hpp:74-83)."""
    _check(library().kx_mixture_avg_transport_props(nStates, offsetT, offset, pressure, _ptr(o_state),
                                                    _ptr(o_viscosity), _ptr(o_thermalConductivity),
                                                    _ptr(o_densityDiffCoeffs), dtype, _stream(stream)),
           'kinetix.mixtureAvgTransportProps')


def thermodynamicProps(n_states, offsetT, offset, pressure, o_state, o_rho, o_cpi, o_rhoCp, stream=None,
                       dtype=KX_DTYPE_F64):
    """kinetix::thermodynamicProps (kinetix.hpp:85-94)."""
    _check(library().kx_thermodynamic_props(n_states, offsetT, offset, pressure, _ptr(o_state), _ptr(o_rho),
                                            _ptr(o_cpi), _ptr(o_rhoCp), dtype, _stream(stream)),
           'kinetix.thermodynamicProps')


def productionRatesPressureField(n_states, offsetT, offset, o_pressure, o_state, o_rates, stream=None,
                                 dtype=KX_DTYPE_F64):
    """extension: productionRates with one pressure per state, o_pressure[id] = p / p_ref."""
    _check(library().kx_production_rates_pfield(n_states, offsetT, offset, _ptr(o_pressure), _ptr(o_state),
                                                _ptr(o_rates), dtype, _stream(stream)),
           'kinetix.productionRatesPressureField')


def thermodynamicPropsPressureField(n_states, offsetT, offset, o_pressure, o_state, o_rho, o_cpi, o_rhoCp,
                                    stream=None, dtype=KX_DTYPE_F64):
    """extension: thermodynamicProps with one pressure per state."""
    _check(library().kx_thermodynamic_props_pfield(n_states, offsetT, offset, _ptr(o_pressure), _ptr(o_state),
                                                   _ptr(o_rho), _ptr(o_cpi), _ptr(o_rhoCp), dtype, _stream(stream)),
           'kinetix.thermodynamicPropsPressureField')


def productionRatesHost(n_states, offsetT, offset, pressure, h_state, h_rates):
    """Host-buffer variant (numpy arrays / pinned torch CPU tensors): H2D + kernel + D2H, pipelined."""
    _check(library().kx_production_rates_host(n_states, offsetT, offset, pressure, _ptr(h_state), _ptr(h_rates)),
           'kinetix.productionRatesHost')


def mixtureAvgTransportPropsHost(nStates, offsetT, offset, pressure, h_state, h_viscosity, h_conductivity, h_rhoD):
    _check(library().kx_mixture_avg_transport_props_host(nStates, offsetT, offset, pressure, _ptr(h_state),
                                                         _ptr(h_viscosity), _ptr(h_conductivity), _ptr(h_rhoD)),
           'kinetix.mixtureAvgTransportPropsHost')


def ratesAndTransportHost(n_states, offsetT, offset, pressure, h_state, h_rates, h_viscosity, h_conductivity, h_rhoD):
    """productionRates + mixtureAvgTransportProps on host buffers with one upload of the state slab."""
    _check(library().kx_rates_and_transport_host(n_states, offsetT, offset, pressure, _ptr(h_state), _ptr(h_rates),
                                                 _ptr(h_viscosity), _ptr(h_conductivity), _ptr(h_rhoD)),
           'kinetix.ratesAndTransportHost')


def nSpecies():
    return library().kx_n_species()


def nActiveSpecies():
    return library().kx_n_active_species()


def nReactions():
    return library().kx_n_reactions()


def speciesNames():
    lib = library()
    return [lib.kx_species_name(k).decode() for k in range(lib.kx_n_species())]


def speciesIndex(name):
    return library().kx_species_index(name.encode())


def _vec(fn, what):
    n = nSpecies()
    arr = (_dbl * n)()
    _check(fn(arr), what)
    return list(arr)


def molecularWeights():
    """M_k / Mbar_ref (kinetix.cpp:888-895)."""
    return _vec(library().kx_molecular_weights, 'kinetix.molecularWeights')


def molarMasses():
    return _vec(library().kx_molar_masses, 'kinetix.molarMasses')


def refPressure():
    return library().kx_ref_pressure()


def refTemperature():
    return library().kx_ref_temperature()


def refMassFractions():
    return _vec(library().kx_ref_mass_fractions, 'kinetix.refMassFractions')


def refMeanMolecularWeight():
    return library().kx_ref_mean_molecular_weight()


def modulePath():
    return library().kx_module_path().decode()


def finalize():
    library().kx_finalize()
