"""ctypes binding of libchroma_b200.so (the C ABI declared in include/chroma_b200.h).

This replaces PyCUDA's role in the reference (compile/allocate/launch:
chroma/gpu/tools.py:45-63, 207-229).  There is NO fallback: if the shared
library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CHROMA_B200_LIB', os.path.join(_HERE, 'libchroma_b200.so'))   # override: A/B builds

u64, i32, u32, f32, vp = C.c_uint64, C.c_int32, C.c_uint32, C.c_float, C.c_void_p


class CbMaterial(C.Structure):
    _fields_ = [(n, i32) for n in (
        'refractive_index', 'absorption_length', 'scattering_length', 'num_comp',
        'comp_reemission_prob', 'comp_reemission_wvl_cdf', 'comp_reemission_time_cdf',
        'comp_absorption_length')]


class CbSurface(C.Structure):
    _fields_ = ([(n, i32) for n in (
        'detect', 'absorb', 'reemit', 'reflect_diffuse', 'reflect_specular', 'eta', 'k',
        'reemission_cdf', 'model', 'transmissive')] + [('thickness', f32)] +
        [(n, i32) for n in (
            'dichroic_nangles', 'dichroic_angles', 'dichroic_reflect', 'dichroic_transmit',
            'angular_nangles', 'angular_angles', 'angular_transmit',
            'angular_reflect_specular', 'angular_reflect_diffuse')])


class CbWirePlane(C.Structure):
    _fields_ = [('origin', f32 * 3), ('u', f32 * 3), ('v', f32 * 3),
                ('pitch', f32), ('radius', f32), ('umin', f32), ('umax', f32), ('vmin', f32), ('vmax', f32), ('v0', f32),
                ('surface_index', i32), ('material_outer_index', i32), ('material_inner_index', i32), ('color', u32)]


class CbGeometryDesc(C.Structure):
    _fields_ = [
        ('vertices', vp), ('nvertices', u64),
        ('triangles', vp), ('ntriangles', u64),
        ('material_codes', vp), ('solid_id', vp), ('colors', vp),
        ('nodes', vp), ('nnodes', u64),
        ('world_origin', f32 * 3), ('world_scale', f32),
        ('table_pool', vp), ('table_floats', u64),
        ('materials', C.POINTER(CbMaterial)), ('nmaterials', i32),
        ('surfaces', C.POINTER(CbSurface)), ('nsurfaces', i32),
        ('wavelength_n', i32), ('wavelength_start', f32), ('wavelength_step', f32),
        ('time_n', i32), ('time_start', f32), ('time_step', f32),
        ('nwireplanes', i32), ('wireplanes', C.POINTER(CbWirePlane)),
    ]


class CbGeometryInfo(C.Structure):
    _fields_ = ([(n, vp) for n in (
        'vertices', 'triangles', 'material_codes', 'colors', 'solid_id_map', 'nodes',
        'solid_id_to_channel_index', 'time_cdf_x', 'time_cdf_y', 'charge_cdf_x', 'charge_cdf_y')] +
        [('nvertices', u64), ('ntriangles', u64), ('nnodes', u64), ('nchannels', i32),
         ('device_bytes', u64), ('max_stack_depth', u32)])


class CbPhotonBank(C.Structure):
    _fields_ = [(n, vp) for n in ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles',
                                  'flags', 'weights', 'evidx')] + [('n', u64)]


class CbPropagateStats(C.Structure):
    _fields_ = [('photons', u64), ('steps', u64), ('nodes_visited', u64), ('tris_tested', u64), ('rays_resolved', u64),
                ('launches', u32), ('kernel_ms', f32), ('intersect0_ms', f32), ('intersect0_rays', u64),
                ('intersect_ms', f32), ('physics_ms', f32), ('tail_ms', f32), ('intersect_rays', u64),
                ('physics_steps', u64), ('tail_photons', u64), ('tail_steps', u64)]


# name -> (restype, argtypes); every symbol include/chroma_b200.h declares
_P = C.POINTER
SIGNATURES = {
    'cb_init': (C.c_int, [C.c_int]),
    'cb_device_count': (C.c_int, []),
    'cb_abi_version': (C.c_int, []),
    'cb_last_error': (C.c_char_p, []),
    'cb_synchronize': (C.c_int, []),
    'cb_sm_count': (C.c_int, []),
    'cb_device_pci_bus_id': (C.c_int, [C.c_char_p, i32]),
    'cb_malloc': (C.c_int, [u64, _P(vp)]),
    'cb_free': (C.c_int, [vp]),
    'cb_memcpy_h2d': (C.c_int, [vp, vp, u64]),
    'cb_memcpy_d2h': (C.c_int, [vp, vp, u64]),
    'cb_memcpy_d2d': (C.c_int, [vp, vp, u64]),
    'cb_memset32': (C.c_int, [vp, u32, u64]),
    'cb_host_alloc': (C.c_int, [u64, _P(vp)]),
    'cb_host_alloc_flags': (C.c_int, [u64, i32, _P(vp)]),
    'cb_host_free': (C.c_int, [vp]),
    'cb_mem_info': (C.c_int, [_P(u64), _P(u64)]),
    'cb_timer_start': (C.c_int, []),
    'cb_timer_stop': (C.c_int, [_P(f32)]),
    'cb_flush_l2': (C.c_int, []),
    'cb_geometry_create': (C.c_int, [_P(CbGeometryDesc), _P(u64)]),
    'cb_geometry_destroy': (C.c_int, [u64]),
    'cb_geometry_info': (C.c_int, [u64, _P(CbGeometryInfo)]),
    'cb_detector_attach': (C.c_int, [u64, vp, u64, i32, vp, vp, i32, vp, vp, i32, f32]),
    'cb_bvh_build': (C.c_int, [vp, u64, vp, u64, i32, _P(f32 * 3), _P(f32), vp, _P(u64), vp, _P(i32)]),
    'cb_native_tree_build': (C.c_int, [vp, u64, u64, vp, vp, _P(u64)]),
    'cb_native_tree_build_split': (C.c_int, [vp, u64, u64, vp, vp, vp, vp, f32, i32, i32, f32, vp, _P(u64)]),
    'cb_rng_create': (C.c_int, [u64, u64, u64, _P(u64)]),
    'cb_rng_create_streams': (C.c_int, [u64, u64, u64, u64, _P(u64)]),
    'cb_rng_view': (C.c_int, [u64, u64, u64, _P(u64)]),
    'cb_rng_destroy': (C.c_int, [u64]),
    'cb_rng_size': (C.c_int, [u64, _P(u64)]),
    'cb_rng_download': (C.c_int, [u64, u64, u64, vp]),
    'cb_rng_fill_uniform': (C.c_int, [u64, u64, f32, f32, vp]),
    'cb_intersect': (C.c_int, [u64, vp, vp, vp, u64, vp, vp]),
    'cb_propagate': (C.c_int, [_P(CbPhotonBank), u64, u64, i32, i32, i32, i32, i32, _P(CbPropagateStats)]),
    'cb_photon_bank_upload': (C.c_int, [_P(CbPhotonBank), _P(CbPhotonBank), u64, u32]),
    'cb_photon_duplicate': (C.c_int, [_P(CbPhotonBank), u64, i32]),
    'cb_count_photons': (C.c_int, [_P(CbPhotonBank), u64, u64, u32, _P(u32)]),
    'cb_copy_photons': (C.c_int, [_P(CbPhotonBank), u64, u64, u32, _P(CbPhotonBank), _P(u32)]),
    'cb_count_photon_hits': (C.c_int, [_P(CbPhotonBank), u64, u64, u32, u64, _P(u32)]),
    'cb_copy_photon_hits': (C.c_int, [_P(CbPhotonBank), u64, u64, u32, u64, _P(CbPhotonBank), vp, _P(u32)]),
    'cb_copy_photon_queue': (C.c_int, [_P(CbPhotonBank), vp, u64, _P(CbPhotonBank)]),
    'cb_copy_photon_hits_async': (C.c_int, [_P(CbPhotonBank), u64, u64, u32, u64, vp, vp]),
    'cb_event_create': (C.c_int, [_P(u64)]),
    'cb_event_record': (C.c_int, [u64]),
    'cb_event_wait': (C.c_int, [u64]),
    'cb_event_destroy': (C.c_int, [u64]),
    'cb_daq_create': (C.c_int, [u64, i32, _P(u64)]),
    'cb_daq_destroy': (C.c_int, [u64]),
    'cb_daq_begin_acquire': (C.c_int, [u64]),
    'cb_daq_acquire': (C.c_int, [u64, _P(CbPhotonBank), u64, i32, i32, u64, u64, f32]),
    'cb_daq_end_acquire': (C.c_int, [u64]),
    'cb_daq_acquire_async': (C.c_int, [u64, _P(CbPhotonBank), u64, i32, i32, u64, u64, f32, i32, i32]),
    'cb_daq_pointers': (C.c_int, [u64, _P(vp), _P(vp), _P(vp), _P(vp), _P(vp), _P(u64)]),
    'cb_daq_finalize': (C.c_int, [u64]),
    'cb_daq_fold': (C.c_int, [u64, u64]),
    'cb_daq_fold_async': (C.c_int, [u64, u64]),
    'cb_comm_unique_id': (C.c_int, [vp]),
    'cb_comm_init': (C.c_int, [i32, i32, vp]),
    'cb_comm_destroy': (C.c_int, []),
    'cb_comm_size': (C.c_int, [_P(i32), _P(i32)]),
    'cb_daq_allreduce': (C.c_int, [u64]),
    'cb_daq_reduce_local': (C.c_int, [_P(u64), i32]),
    'cb_set_blocking_sync': (C.c_int, [i32]),
    'cb_unique_vertices': (C.c_int, [vp, u64, vp, vp, _P(u64)]),
    'cb_pdf_bin_hits': (C.c_int, [i32, vp, vp, vp, i32, f32, f32, i32, f32, f32, vp]),
    'cb_pdf_accumulate_moments': (C.c_int, [i32, i32, vp, vp, f32, f32, f32, f32, vp, vp, vp, vp, vp]),
    'cb_pdf_accumulate_kernel_eval': (C.c_int, [i32, i32, vp, vp, vp, vp, vp, f32, f32, f32, f32, vp, vp, vp, vp, vp]),
    'cb_pdf_accumulate_eval': (C.c_int, [i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, f32, i32, vp, vp]),
}

_lib = None
_device = None
numa_cores = None        # cores this process was bound to at init() (parallel.bind_to_gpu_numa_node), or None


class ChromaB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed for this); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'chroma_lite_b200: %s is missing. Build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`make -C chroma_lite_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().cb_last_error()
        raise ChromaB200Error('libchroma_b200 error %d: %s' % (rc, msg.decode() if msg else '?'))


def init(device=None):
    """Bind the library to a CUDA device (one process per GPU)."""
    global _device
    lib = load()
    if device is None:
        if _device is not None:
            return _device
        device = int(os.environ.get('LOCAL_RANK', '0')) if lib.cb_device_count() > 1 else 0
    if _device is not None and _device == device:
        return _device
    if lib.cb_device_count() < 1:
        raise ChromaB200Error('chroma_lite_b200 needs a CUDA device (sm_100a); none is visible. '
                              'There is no CPU fallback.')
    check(lib.cb_init(int(device)))
    _device = int(device)
    # several ranks on one host: stay on the cores (and so the memory) next to this rank's GPU
    if int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')) or 1) > 1 or os.environ.get('CHROMA_B200_NUMA') == '1':
        from . import parallel
        buf = C.create_string_buffer(32)
        if lib.cb_device_pci_bus_id(buf, 32) == 0:
            global numa_cores
            numa_cores = parallel.bind_to_gpu_numa_node(buf.value.decode())
    return _device


def lib():
    """The loaded library, bound to a device."""
    if _device is None:
        init()
    return _lib
