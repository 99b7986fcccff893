"""Host-side geometry input model: Mesh / Solid / Material / Surface / Geometry.

This is the data contract GPUGeometry consumes (flattened arrays, material and
surface property tables); attribute names follow chroma/geometry.py so that the
reference's own objects can be handed to the engine unchanged (duck typing).
Nothing here runs on the GPU.
"""
import numpy as np

# all material/surface properties are resampled onto this grid for the device
# (chroma/geometry.py:15-17): 60..995 nm in 5 nm steps, 188 points
standard_wavelengths = np.arange(60, 1000, 5).astype(np.float32)


NATIVE_UNIQUE_MIN = 200000      # vertices; below this np.unique is quick enough


def _native_unique_vertices(vertices):
    """(unique rows, inverse) from libchroma_b200's host routine, or (None, None) when the
    library cannot be loaded (host-only convenience: NumPy then does the same job)."""
    import ctypes as C
    try:
        from . import _lib
        lib = _lib.load()
    except Exception:
        return None, None
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    n = len(v)
    uniq = np.empty((n, 3), dtype=np.float32)
    inverse = np.empty(n, dtype=np.uint32)
    count = C.c_uint64()
    if lib.cb_unique_vertices(v.ctypes.data, n, uniq.ctypes.data, inverse.ctypes.data, C.byref(count)) != 0:
        return None, None
    return uniq[:count.value].copy(), inverse.astype(np.intp)


class Mesh(object):
    """Triangle mesh: float32 vertices (V,3) and integer triangles (T,3)."""

    def __init__(self, vertices, triangles, remove_duplicate_vertices=False, round=True,
                 remove_null_triangles=True):
        vertices = np.asarray(vertices, dtype=np.float32)
        triangles = np.asarray(triangles, dtype=np.int32)
        if vertices.ndim != 2 or vertices.shape[1] != 3 or triangles.ndim != 2 or triangles.shape[1] != 3:
            raise ValueError('shape mismatch')
        if (triangles < 0).any():
            raise ValueError('indices in `triangles` must be positive.')
        if (triangles >= len(vertices)).any():
            raise ValueError('indices in `triangles` must be less than the length of the vertex array.')
        self.vertices = vertices.round(decimals=12) if round else vertices
        self.triangles = triangles
        if remove_duplicate_vertices:
            self.remove_duplicate_vertices()
        if remove_null_triangles:
            self.remove_null_triangles()

    def md5(self):
        """MD5 of the vertex array, then the triangle array, hex (chroma/geometry.py:107-112)."""
        from .cache import mesh_hash
        return mesh_hash(self)

    def get_bounds(self):
        return np.min(self.vertices, axis=0), np.max(self.vertices, axis=0)

    def get_triangle_centers(self):
        return np.mean(self.assemble(), axis=1)

    def remove_duplicate_vertices(self):
        # unique rows in lexicographic order; triangles remapped through the inverse
        # (chroma/geometry.py:59-69).  Large meshes use the library's multi-threaded host
        # routine (cb_unique_vertices, same result as np.unique; no GPU involved): the NumPy
        # sort of the 29k-PMT detector's 18.6 M vertices takes 15-30 s on one core.
        uniq = inverse = None
        if len(self.vertices) >= NATIVE_UNIQUE_MIN:
            uniq, inverse = _native_unique_vertices(self.vertices)
        if uniq is None:
            uniq, inverse = np.unique(self.vertices, axis=0, return_inverse=True)
        self.vertices = np.ascontiguousarray(uniq)
        self.triangles = np.asarray(inverse).reshape(-1)[self.triangles]

    def remove_null_triangles(self):
        if len(self.triangles) == 0:
            return None
        t = self.triangles
        mask = (t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])
        self.triangles = t[mask]
        return mask

    def assemble(self, key=slice(None), group=True):
        idx = self.triangles[key] if group else self.triangles[key].flatten()
        return self.vertices[idx]

    def __add__(self, other):
        return Mesh(np.concatenate((self.vertices, other.vertices)),
                    np.concatenate((self.triangles, other.triangles + len(self.vertices))))


def _per_triangle(value, n, dtype=object):
    if isinstance(value, (list, tuple, np.ndarray)):
        if len(value) != n:
            raise ValueError('shape mismatch')
        return np.array(value, dtype=dtype)
    out = np.empty(n, dtype=dtype)
    out[:] = value
    return out


def _unique_objects(arr):
    seen, out = set(), []
    for x in arr:
        if id(x) not in seen:
            seen.add(id(x))
            out.append(x)
    return out


class Solid(object):
    """A mesh with inner material (material1), outer material (material2),
    optional surface and colour per triangle (chroma/geometry.py:117-153)."""

    def __init__(self, mesh, material1=None, material2=None, surface=None, color=0x33ffffff):
        n = len(mesh.triangles)
        self.mesh = mesh
        self.material1 = _per_triangle(material1, n)
        self.material2 = _per_triangle(material2, n)
        self.surface = _per_triangle(surface, n)
        self.color = _per_triangle(color, n, dtype=np.uint32)
        self.unique_materials = _unique_objects(np.concatenate([self.material1, self.material2]))
        self.unique_surfaces = _unique_objects(self.surface)

    def __add__(self, other):
        return Solid(self.mesh + other.mesh, np.concatenate((self.material1, other.material1)),
                     np.concatenate((self.material2, other.material2)),
                     np.concatenate((self.surface, other.surface)),
                     np.concatenate((self.color, other.color)))


class Material(object):
    """Bulk optical properties; each property is an (n,2) array of (wavelength, value)."""

    def __init__(self, name='none'):
        self.name = name
        self.refractive_index = None
        self.absorption_length = None
        self.scattering_length = None
        self.comp_reemission_prob = []
        self.comp_reemission_wvl_cdf = []
        self.comp_reemission_times = []
        self.comp_reemission_time_cdf = []
        self.comp_absorption_length = []
        self.density = 0.0
        self.composition = {}

    def set(self, name, value, wavelengths=standard_wavelengths):
        if np.iterable(value):
            if len(value) != len(wavelengths):
                raise ValueError('shape mismatch')
        else:
            value = np.tile(value, len(wavelengths))
        self.__dict__[name] = np.array(list(zip(wavelengths, value)), dtype=np.float32)

    def __repr__(self):
        return '<Material %s>' % self.name


vacuum = Material('vacuum')
vacuum.set('refractive_index', 1.0)
vacuum.set('absorption_length', 1e6)
vacuum.set('scattering_length', 1e6)


class DichroicProps(object):
    def __init__(self, angles, reflect, transmit):
        self.angles = np.asarray(angles)
        self.dichroic_reflect = np.asarray(reflect)      # [angle][(wavelength, value)]
        self.dichroic_transmit = np.asarray(transmit)


class AngularProps(object):
    def __init__(self, angles, transmit, reflect_specular=None, reflect_diffuse=None):
        self.angles = np.asarray(angles)
        self.transmit = np.asarray(transmit)
        self.reflect_specular = np.zeros_like(self.transmit) if reflect_specular is None else np.asarray(reflect_specular)
        self.reflect_diffuse = np.zeros_like(self.transmit) if reflect_diffuse is None else np.asarray(reflect_diffuse)


SURFACE_DEFAULT, SURFACE_COMPLEX, SURFACE_WLS, SURFACE_DICHROIC, SURFACE_ANGULAR = range(5)


class Surface(object):
    """Surface optical properties (chroma/geometry.py:262-295)."""

    def __init__(self, name='none', model=0):
        self.name = name
        self.model = model
        for prop in ('detect', 'absorb', 'reemit', 'reflect_diffuse', 'reflect_specular', 'eta', 'k',
                     'reemission_cdf'):
            self.set(prop, 0)
        self.dichroic_props = None
        self.angular_props = None
        self.thickness = 0.0
        self.transmissive = 0

    def set(self, name, value, wavelengths=standard_wavelengths):
        if np.iterable(value):
            if len(value) != len(wavelengths):
                raise ValueError('shape mismatch')
        else:
            value = np.tile(value, len(wavelengths))
        if (np.asarray(value) < 0.0).any():
            raise Exception('all probabilities must be >= 0.0')
        self.__dict__[name] = np.array(list(zip(wavelengths, value)), dtype=np.float32)

    def __repr__(self):
        return '<Surface %s>' % self.name


class Geometry(object):
    """A list of placed solids that flattens into one triangle soup
    (chroma/geometry.py:297-391)."""

    def __init__(self, detector_material=None):
        self.detector_material = detector_material
        self.solids = []
        self.solid_rotations = []
        self.solid_displacements = []
        self.bvh = None

    def add_solid(self, solid, rotation=None, displacement=None):
        rotation = np.identity(3, dtype=np.float32) if rotation is None else np.asarray(rotation, dtype=np.float32)
        if rotation.shape != (3, 3):
            raise ValueError('rotation matrix has the wrong shape.')
        displacement = np.zeros(3, dtype=np.float32) if displacement is None else np.asarray(displacement, dtype=np.float32)
        if displacement.shape != (3,):
            raise ValueError('displacement vector has the wrong shape.')
        self.solid_rotations.append(rotation)
        self.solid_displacements.append(displacement)
        self.solids.append(solid)
        return len(self.solids) - 1

    def flatten(self, dedupe_vertices=True):
        if hasattr(self, 'mesh'):
            return
        nv = np.cumsum([0] + [len(s.mesh.vertices) for s in self.solids])
        nt = np.cumsum([0] + [len(s.mesh.triangles) for s in self.solids])
        vertices = np.empty((nv[-1], 3), dtype=np.float32)
        triangles = np.empty((nt[-1], 3), dtype=np.uint32)
        for i, solid in enumerate(self.solids):
            vertices[nv[i]:nv[i + 1]] = np.inner(solid.mesh.vertices, self.solid_rotations[i]) + self.solid_displacements[i]
            triangles[nt[i]:nt[i + 1]] = solid.mesh.triangles + nv[i]
        self.mesh = Mesh(vertices, triangles, remove_duplicate_vertices=dedupe_vertices,
                         remove_null_triangles=False)
        self.colors = np.concatenate([s.color for s in self.solids])
        self.solid_id = np.concatenate([np.full(len(s.mesh.triangles), i, dtype=np.uint32)
                                        for i, s in enumerate(self.solids)])
        self.unique_materials = _unique_objects([m for s in self.solids for m in s.unique_materials])
        mat_lookup = {id(m): i for i, m in enumerate(self.unique_materials)}
        self.unique_surfaces = _unique_objects([x for s in self.solids for x in s.unique_surfaces])
        surf_lookup = {id(x): i for i, x in enumerate(self.unique_surfaces)}
        cache = {}

        def indices(solid, attr, lookup):
            key = (id(solid), attr)
            if key not in cache:
                cache[key] = np.fromiter((lookup[id(x)] for x in getattr(solid, attr)), dtype=np.int32,
                                         count=len(solid.mesh.triangles))
            return cache[key]

        self.material1_index = np.concatenate([indices(s, 'material1', mat_lookup) for s in self.solids])
        self.material2_index = np.concatenate([indices(s, 'material2', mat_lookup) for s in self.solids])
        self.surface_index = np.concatenate([indices(s, 'surface', surf_lookup) for s in self.solids])
        if id(None) in surf_lookup:
            self.surface_index[self.surface_index == surf_lookup[id(None)]] = -1

    def build(self, target_degree=3):
        """flatten() + recursive-grid BVH (needs the GPU library for the build)."""
        from .bvh import make_recursive_grid_bvh
        self.flatten()
        if self.bvh is None:
            self.bvh = make_recursive_grid_bvh(self.mesh, target_degree=target_degree)
        return self
