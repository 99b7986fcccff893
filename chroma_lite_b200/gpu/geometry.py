"""GPUGeometry: resample the optical tables, pack them into one pool and hand the
flattened mesh + BVH to the engine (role of chroma/gpu/geometry.py:14-526)."""
import ctypes as C
import numpy as np

from .. import _lib
from ..gpuarray import DeviceArray, vec
from ..geometry import standard_wavelengths
from .tools import format_array, format_size


def _uniform_step(x, what):
    d = np.unique(np.diff(x))
    if len(d) != 1:
        raise ValueError('%s must be equally spaced apart.' % what)
    return float(d.item())


class _Pool(object):
    """Float pool builder: wavelength tables first, long time CDFs last (the
    engine stages the leading part in shared memory)."""

    def __init__(self):
        self.front, self.back = [], []
        self.nfront = 0

    def add(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32).ravel()
        off = self.nfront
        self.front.append(arr)
        self.nfront += len(arr)
        return off

    def add_back(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32).ravel()
        self.back.append(arr)
        return len(self.back) - 1          # resolved in finish()

    def finish(self):
        offs, cur = [], self.nfront
        for a in self.back:
            offs.append(cur)
            cur += len(a)
        pool = np.concatenate(self.front + self.back) if (self.front or self.back) else np.zeros(0, np.float32)
        return pool.astype(np.float32), offs


def interp_property(wavelengths, prop):
    """Linear resampling onto the device grid (chroma/gpu/geometry.py:44-49)."""
    assert prop is not None, 'property must not be None'
    prop = np.asarray(prop)
    return np.interp(wavelengths, prop[:, 0], prop[:, 1]).astype(np.float32)


def wireplane_lists(geometry):
    """Material and surface lists extended by the objects that only analytic wire planes
    reference (chroma/gpu/geometry.py:109-112, 265-270), plus the plane descriptors."""
    materials = list(geometry.unique_materials)
    surfaces = list(geometry.unique_surfaces)
    planes = list(getattr(geometry, 'wireplanes', None) or [])
    for desc in planes:
        for mat in (desc.get('material_inner', None), desc.get('material_outer', None)):
            if mat is not None and mat not in materials:
                materials.append(mat)
    for desc in planes:
        surface = desc.get('surface', None)
        if surface is not None and surface not in surfaces:
            surfaces.append(surface)
    return materials, surfaces, planes


def build_wireplanes(planes, materials, surfaces):
    """CbWirePlane[] from the reference's dict descriptors (chroma/gpu/geometry.py:343-387)."""
    arr = (_lib.CbWirePlane * max(1, len(planes)))()
    for i, desc in enumerate(planes):
        wp = arr[i]
        for name in ('origin', 'u', 'v'):
            vals = np.asarray(desc[name], dtype=np.float32)
            setattr(wp, name, (C.c_float * 3)(*[float(x) for x in vals]))
        for name in ('pitch', 'radius', 'umin', 'umax', 'vmin', 'vmax', 'v0'):
            setattr(wp, name, float(np.float32(desc[name])))
        surface = desc.get('surface', None)
        wp.surface_index = -1 if surface is None else (surfaces.index(surface) if surface in surfaces else -1)
        mo, mi = desc.get('material_outer', None), desc.get('material_inner', None)
        if mo is None or mi is None:           # "should not happen under normal use": the reference falls back to slot 0
            wp.material_outer_index = wp.material_inner_index = 0
        else:
            wp.material_outer_index, wp.material_inner_index = materials.index(mo), materials.index(mi)
        wp.color = int(desc.get('color', 0)) & 0xFFFFFFFF
    return arr


def build_tables(geometry, wavelengths, times, materials=None, surfaces=None):
    """(pool, materials[], surfaces[]) for CbGeometryDesc from a flattened geometry."""
    pool = _Pool()
    materials = list(geometry.unique_materials) if materials is None else materials
    surfaces = list(geometry.unique_surfaces) if surfaces is None else surfaces
    mats = (_lib.CbMaterial * max(1, len(materials)))()
    pending_time = []
    W = len(wavelengths)
    for i, m in enumerate(materials):
        if m is None:
            raise Exception('one or more triangles is missing a material.')
        cm = mats[i]
        cm.refractive_index = pool.add(interp_property(wavelengths, m.refractive_index))
        cm.absorption_length = pool.add(interp_property(wavelengths, m.absorption_length))
        cm.scattering_length = pool.add(interp_property(wavelengths, m.scattering_length))
        ncomp = len(getattr(m, 'comp_reemission_prob', []))
        for name in ('comp_reemission_wvl_cdf', 'comp_reemission_time_cdf', 'comp_absorption_length'):
            assert ncomp == len(getattr(m, name, [])), 'component arrays must be same length'
        cm.num_comp = ncomp
        cm.comp_reemission_prob = cm.comp_reemission_wvl_cdf = -1
        cm.comp_reemission_time_cdf = cm.comp_absorption_length = -1
        if ncomp:
            cm.comp_reemission_prob = pool.add(np.concatenate([interp_property(wavelengths, c) for c in m.comp_reemission_prob]))
            cm.comp_reemission_wvl_cdf = pool.add(np.concatenate([interp_property(wavelengths, c) for c in m.comp_reemission_wvl_cdf]))
            cm.comp_absorption_length = pool.add(np.concatenate([interp_property(wavelengths, c) for c in m.comp_absorption_length]))
            pending_time.append((i, pool.add_back(np.concatenate([interp_property(times, c) for c in m.comp_reemission_time_cdf]))))
    surfs = (_lib.CbSurface * max(1, len(surfaces)))()
    for i, s in enumerate(surfaces):
        cs = surfs[i]
        for f, _t in _lib.CbSurface._fields_:
            if f != 'thickness':
                setattr(cs, f, -1)
        cs.thickness = 0.0
        cs.transmissive = 0
        cs.dichroic_nangles = cs.angular_nangles = 0
        if s is None:
            continue                      # null slot, never referenced (surface index -1)
        for name in ('detect', 'absorb', 'reemit', 'reflect_diffuse', 'reflect_specular', 'eta', 'k', 'reemission_cdf'):
            setattr(cs, name, pool.add(interp_property(wavelengths, getattr(s, name))))
        cs.model = int(s.model)
        cs.transmissive = int(s.transmissive)
        cs.thickness = float(s.thickness)
        dp = getattr(s, 'dichroic_props', None)
        if dp:
            cs.dichroic_nangles = len(dp.angles)
            cs.dichroic_angles = pool.add(np.asarray(dp.angles, dtype=np.float32))
            cs.dichroic_reflect = pool.add(np.concatenate([interp_property(wavelengths, r) for r in dp.dichroic_reflect]))
            cs.dichroic_transmit = pool.add(np.concatenate([interp_property(wavelengths, r) for r in dp.dichroic_transmit]))
        ap = getattr(s, 'angular_props', None)
        if ap:
            cs.angular_nangles = len(ap.angles)
            cs.angular_angles = pool.add(np.asarray(ap.angles, dtype=np.float32))
            cs.angular_transmit = pool.add(np.asarray(ap.transmit, dtype=np.float32))
            cs.angular_reflect_specular = pool.add(np.asarray(ap.reflect_specular, dtype=np.float32))
            cs.angular_reflect_diffuse = pool.add(np.asarray(ap.reflect_diffuse, dtype=np.float32))
    pool_arr, back_offs = pool.finish()
    for i, k in pending_time:
        mats[i].comp_reemission_time_cdf = back_offs[k]
    assert W == len(wavelengths)
    return pool_arr, mats, surfs


def material_codes(geometry):
    """m1<<24 | m2<<16 | surface<<8 with 8-bit two's-complement fields
    (chroma/gpu/geometry.py:401-403)."""
    return (((np.asarray(geometry.material1_index) & 0xff) << 24) |
            ((np.asarray(geometry.material2_index) & 0xff) << 16) |
            ((np.asarray(geometry.surface_index) & 0xff) << 8)).astype(np.uint32)


def make_desc(geometry, wavelengths=None, times=None):
    """Build the CbGeometryDesc (plus the numpy arrays that must stay alive)."""
    if wavelengths is None:
        wavelengths = standard_wavelengths
    wavelengths = np.asarray(wavelengths, dtype=np.float32)
    wavelength_step = _uniform_step(wavelengths, 'wavelengths')
    if times is None:
        time_step = 0.05
        times = np.arange(0, 1000, time_step)
    else:
        time_step = _uniform_step(times, 'times')
    if not hasattr(geometry, 'mesh'):
        geometry.flatten()
    if geometry.bvh is None:
        from ..bvh import make_recursive_grid_bvh
        geometry.bvh = make_recursive_grid_bvh(geometry.mesh)
    keep = {}
    keep['vertices'] = np.ascontiguousarray(geometry.mesh.vertices, dtype=np.float32)
    keep['triangles'] = np.ascontiguousarray(geometry.mesh.triangles, dtype=np.uint32)
    keep['codes'] = material_codes(geometry)
    keep['solid_id'] = np.ascontiguousarray(geometry.solid_id, dtype=np.uint32)
    keep['colors'] = np.ascontiguousarray(geometry.colors, dtype=np.uint32)
    nodes = np.ascontiguousarray(geometry.bvh.nodes)
    keep['nodes'] = nodes
    materials, surfaces, planes = wireplane_lists(geometry)
    pool, mats, surfs = build_tables(geometry, wavelengths, times, materials, surfaces)
    keep['pool'], keep['mats'], keep['surfs'] = pool, mats, surfs
    keep['wireplanes'] = build_wireplanes(planes, materials, surfaces)
    d = _lib.CbGeometryDesc()
    d.vertices, d.nvertices = keep['vertices'].ctypes.data, len(keep['vertices'])
    d.triangles, d.ntriangles = keep['triangles'].ctypes.data, len(keep['triangles'])
    d.material_codes = keep['codes'].ctypes.data
    d.solid_id = keep['solid_id'].ctypes.data
    d.colors = keep['colors'].ctypes.data
    d.nodes, d.nnodes = nodes.ctypes.data, len(nodes)
    wo = np.asarray(geometry.bvh.world_coords.world_origin, dtype=np.float32)
    d.world_origin = (C.c_float * 3)(*[float(x) for x in wo])
    d.world_scale = float(np.float32(geometry.bvh.world_coords.world_scale))
    d.table_pool, d.table_floats = pool.ctypes.data, len(pool)
    d.materials, d.nmaterials = mats, len(materials)
    d.surfaces, d.nsurfaces = surfs, len(surfaces)
    d.wavelength_n, d.wavelength_start, d.wavelength_step = len(wavelengths), float(wavelengths[0]), wavelength_step
    d.time_n, d.time_start, d.time_step = len(times), float(times[0]), time_step
    d.nwireplanes = len(planes)
    d.wireplanes = keep['wireplanes'] if planes else None
    return d, keep


class GPUGeometry(object):
    def __init__(self, geometry, wavelengths=None, times=None, print_usage=False, min_free_gpu_mem=300e6):
        lib = _lib.lib()
        desc, keep = make_desc(geometry, wavelengths, times)
        h = C.c_uint64()
        _lib.check(lib.cb_geometry_create(C.byref(desc), C.byref(h)))
        self.handle = h.value
        self.gpudata = self.handle          # what kernels receive in the reference; here the engine handle
        self.geometry = geometry
        self._refresh_views()
        self.world_origin = vec.make_float3(*desc.world_origin)
        self.world_scale = np.float32(desc.world_scale)
        if print_usage:
            self.print_device_usage()

    def _refresh_views(self):
        info = _lib.CbGeometryInfo()
        _lib.check(_lib.lib().cb_geometry_info(self.handle, C.byref(info)))
        self._info = info
        view = lambda p, n, dt: DeviceArray(n, dt, _alloc=self, _ptr=p) if p else None
        self.vertices = view(info.vertices, info.nvertices, vec.float3)
        self.triangles = view(info.triangles, info.ntriangles, vec.uint3)
        self.material_codes = view(info.material_codes, info.ntriangles, np.uint32)
        self.colors = view(info.colors, info.ntriangles, np.uint32)
        self.solid_id_map = view(info.solid_id_map, info.ntriangles, np.uint32)
        self.nodes = view(info.nodes, info.nnodes, vec.uint4)
        self.extra_nodes = None             # everything is resident: no host-mapped split on a 180 GB part
        self.device_bytes = info.device_bytes

    def device_usage_str(self):
        s = 'device usage:\n' + '-' * 10 + '\n'
        s += format_array('nodes', self.nodes) + '\n'
        s += '%-15s %6s %6s' % ('total', '', format_size(self.device_bytes)) + '\n' + '-' * 10 + '\n'
        free, total = C.c_uint64(), C.c_uint64()
        _lib.check(_lib.lib().cb_mem_info(C.byref(free), C.byref(total)))
        s += '%-15s %6s %6s' % ('device total', '', format_size(total.value)) + '\n'
        s += '%-15s %6s %6s' % ('device used', '', format_size(total.value - free.value)) + '\n'
        s += '%-15s %6s %6s' % ('device free', '', format_size(free.value)) + '\n'
        return s

    def print_device_usage(self):
        print(self.device_usage_str())
        print()

    def reset_colors(self):
        self.colors.set(np.asarray(self.geometry.colors, dtype=np.uint32))

    def color_solids(self, solid_hit, colors, nblocks_per_thread=64, max_blocks=1024):
        """Recolour the triangles of hit solids (viewer helper, mesh.h:161-175)."""
        solid_hit = np.asarray(solid_hit, dtype=bool)
        colors = np.asarray(colors, dtype=np.uint32)
        sid = np.asarray(self.geometry.solid_id)
        cur = self.colors.get()
        m = solid_hit[sid]
        cur[m] = colors[sid[m]]
        self.colors.set(cur)

    def __del__(self):
        try:
            if getattr(self, 'handle', 0) and _lib._lib is not None:
                _lib._lib.cb_geometry_destroy(self.handle)
        except Exception:
            pass
        self.handle = 0
