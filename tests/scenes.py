"""Shared scene / photon builders for the tests (CPU only; BVH from the oracle
builder so no GPU is needed to construct inputs)."""
import numpy as np

from chroma_lite_b200 import demo, event
from chroma_lite_b200.geometry import Geometry, Solid, Mesh, Material, Surface, DichroicProps, AngularProps, \
    standard_wavelengths, vacuum, SURFACE_COMPLEX, SURFACE_WLS, SURFACE_DICHROIC, SURFACE_ANGULAR
from chroma_lite_b200.detector import Detector
from chroma_lite_b200.make import sphere, box, cube
from chroma_lite_b200.sample import uniform_sphere
from chroma_lite_b200.demo import optics
from chroma_lite_b200.gpu.geometry import make_desc


def with_bvh(geo):
    from oracle import bvh_oracle
    geo.flatten()
    if geo.bvh is None:
        bvh_oracle.attach_bvh(geo)
    return geo


def sphere_scene(nsteps=32):
    return with_bvh(demo.acrylic_sphere_scene(nsteps))


def point_source(n, seed=0, wavelength=400.0, pos=(0, 0, 0), wl_range=None):
    rng = np.random.default_rng(seed)
    d = uniform_sphere(n, rng=rng)
    pol = np.cross(d, uniform_sphere(n, rng=rng))
    pol /= np.linalg.norm(pol, axis=1)[:, None]
    wl = np.full(n, wavelength, dtype=np.float32) if wl_range is None else rng.uniform(wl_range[0], wl_range[1], n)
    return event.Photons(np.tile(np.asarray(pos, dtype=np.float32), (n, 1)), d, pol, wl)


def tiny_detector(pmt_nsteps=6):
    return with_bvh(demo.tiny(pmt_nsteps=pmt_nsteps))


def water_box(size=100.0):
    """Photons inside a water cube with a black outside (test_rayleigh.py setup)."""
    geo = Geometry(optics.water)
    geo.add_solid(Solid(cube(size), optics.water, vacuum, surface=optics.black_surface))
    return with_bvh(geo)


def scintillator_scene(nsteps=24):
    """BASELINE config 4 in miniature: re-emitting scintillator in an acrylic
    vessel, WLS-coated shell, dichroic + angular + thin-film surfaces."""
    wl = standard_wavelengths.astype(np.float64)
    scint = Material('scint')
    scint.set('refractive_index', 1.5)
    scint.set('absorption_length', 800.0 + 4.0 * (wl - 60.0))
    scint.set('scattering_length', 3000.0)
    for k, (mu, frac) in enumerate(((430.0, 0.7), (480.0, 0.3))):
        pdf = np.exp(-0.5 * ((wl - mu) / 25.0) ** 2)
        cdf = np.cumsum(pdf) / np.sum(pdf)
        cdf[0] = 0.0
        scint.comp_reemission_prob.append(np.column_stack([wl, np.where(wl < mu, 0.8, 0.0)]).astype(np.float32))
        scint.comp_reemission_wvl_cdf.append(np.column_stack([wl, cdf]).astype(np.float32))
        tt = np.arange(0, 1000, 0.05)
        tcdf = 1.0 - np.exp(-tt / (5.0 + 20.0 * k))
        tcdf[-1] = 1.0
        scint.comp_reemission_times.append(tt)
        scint.comp_reemission_time_cdf.append(np.column_stack([tt, tcdf]).astype(np.float32))
        scint.comp_absorption_length.append(np.column_stack([wl, (800.0 + 4.0 * (wl - 60.0)) / frac]).astype(np.float32))

    wls = Surface('wls', model=SURFACE_WLS)
    wls.set('absorb', np.where(wl < 420.0, 0.6, 0.05))
    wls.set('reemit', 0.9)
    wls.set('reflect_diffuse', 0.1)
    wls.set('reflect_specular', 0.1)
    pdf = np.exp(-0.5 * ((wl - 500.0) / 30.0) ** 2)
    cdf = np.cumsum(pdf) / np.sum(pdf)
    cdf[0] = 0.0
    wls.set('reemission_cdf', cdf)

    film = Surface('film', model=SURFACE_COMPLEX)
    film.set('detect', 0.3)
    film.set('reflect_diffuse', 0.2)
    film.set('eta', 2.0 + 0.001 * (wl - 400.0).clip(0))
    film.set('k', 1.0)
    film.thickness = 25.0
    film.transmissive = 1

    dich = Surface('dichroic', model=SURFACE_DICHROIC)
    angles = np.array([0.0, 0.4, 0.9, 1.5708])
    refl = [np.column_stack([wl, np.clip((wl - 380.0 - 40.0 * a) / 100.0, 0, 0.9)]) for a in range(4)]
    tran = [np.column_stack([wl, np.clip(0.95 - r[:, 1], 0, 1)]) for r in refl]
    dich.dichroic_props = DichroicProps(angles, refl, tran)

    ang = Surface('angular', model=SURFACE_ANGULAR)
    ang.angular_props = AngularProps(np.linspace(0, np.pi / 2, 7), np.linspace(0.8, 0.1, 7), np.linspace(0.1, 0.5, 7),
                                     np.linspace(0.05, 0.2, 7))

    geo = Detector(optics.water)
    geo.add_solid(Solid(sphere(600.0, nsteps), scint, optics.acrylic))
    geo.add_solid(Solid(sphere(650.0, nsteps), optics.acrylic, optics.water, surface=dich))
    geo.add_solid(Solid(sphere(1200.0, nsteps), optics.water, optics.water, surface=wls))
    geo.add_solid(Solid(box(300, 300, 10, center=(0, 0, 900)), optics.glass, optics.water, surface=ang))
    geo.add_pmt(Solid(box(400, 400, 20, center=(0, 0, -900)), optics.glass, optics.water, surface=film))
    geo.add_pmt(Solid(box(400, 20, 400, center=(0, 900, 0)), optics.glass, optics.water, surface=optics.photocathode))
    geo.add_solid(Solid(sphere(2000.0, nsteps), optics.water, optics.water, surface=optics.black_surface))
    geo.set_time_dist_gaussian(1.2, -6.0, 6.0)
    geo.set_charge_dist_gaussian(1.0, 0.1, 0.5, 1.5)
    return with_bvh(geo)


def wireplane_scene(size=400.0):
    """Water cube with two crossed analytic wire planes (this fork's WirePlane primitive):
    steel-like wires with a half-absorbing, half-reflecting surface, one plane along x at
    z = 0 and one along y at z = 60, one of them referencing a material and a surface that
    no triangle uses (chroma/gpu/geometry.py:109-112, 265-270)."""
    geo = Geometry(optics.water)
    geo.add_solid(Solid(cube(size), optics.water, vacuum, surface=optics.black_surface))
    with_bvh(geo)
    wl = standard_wavelengths.astype(np.float64)
    steel = Material('steel')
    steel.set('refractive_index', 2.5)
    steel.set('absorption_length', 1e-3)
    steel.set('scattering_length', 1e6)
    wire_surface = Surface('wire')
    wire_surface.set('absorb', 0.4)
    wire_surface.set('reflect_specular', 0.3)
    wire_surface.set('reflect_diffuse', 0.3)
    glassy = Surface('glassy')          # transparent wires: exercises the inside branch
    glassy.set('absorb', 0.0)
    h = size / 2 - 20.0
    geo.wireplanes = [
        dict(origin=(0.0, 0.0, 0.0), u=(1.0, 0.0, 0.0), v=(0.0, 1.0, 0.0), pitch=5.0, radius=0.3, umin=-h, umax=h,
             vmin=-h, vmax=h, v0=0.0, surface=wire_surface, material_inner=steel, material_outer=optics.water, color=0x33),
        dict(origin=(0.0, 0.0, 60.0), u=(0.0, 2.0, 0.0), v=(1.0, 0.3, 0.0), pitch=3.0, radius=0.8, umin=-h, umax=h,
             vmin=-0.5 * h, vmax=0.7 * h, v0=0.4, surface=glassy, material_inner=optics.glass, material_outer=optics.water),
    ]
    return geo


def desc_of(geo):
    return make_desc(geo)


def ref_tiny_detector():
    """The reference's demo detector in miniature, built from the reference's own PMT model and
    optics tables (tests/golden/ref_detector_parts.npz, chroma_lite_b200/demo/refparts.py)."""
    from chroma_lite_b200.demo import refparts
    return with_bvh(refparts.tiny())


def many_tables_scene(nsteps=16, nangles=14):
    """More wavelength tables than fit the 48 KB staged into shared memory (the cut must
    fall on a table boundary): the scintillator scene's media plus a dichroic filter with `nangles`
    angles (2 x nangles x 188 floats) placed so that it is hit often."""
    geo = scintillator_scene(nsteps)
    wl = standard_wavelengths.astype(np.float64)
    angles = np.linspace(0.0, np.pi / 2, nangles)
    refl = [np.column_stack([wl, np.clip((wl - 330.0 - 15.0 * a) / 160.0, 0.05, 0.9)]) for a in range(nangles)]
    tran = [np.column_stack([wl, np.clip(0.97 - r[:, 1], 0, 1)]) for r in refl]
    for s in geo.solids:
        for x in s.unique_surfaces:
            if x is not None and x.name == 'dichroic':
                x.dichroic_props = DichroicProps(angles, refl, tran)
    return geo
