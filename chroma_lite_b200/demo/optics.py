"""Optical tables for the demo detectors (role of chroma/demo/optics.py).

The numbers are smooth analytic stand-ins for a water-Cherenkov detector, not
measured data: what matters for the engine is that every table varies with
wavelength so the interpolation paths are exercised."""
import numpy as np

from ..geometry import Material, Surface, standard_wavelengths, vacuum  # noqa: F401

_wl = standard_wavelengths.astype(np.float64)

water = Material('water')
water.set('refractive_index', 1.3247 + 3.3e3 / _wl ** 2)
# absorption length [mm]: clearest in the blue, strong red/IR and UV absorption
_abs_per_m = 0.0045 + 0.25 * np.exp((_wl - 600.0) / 45.0) + 0.4 * np.exp(-(_wl - 200.0) / 40.0)
water.set('absorption_length', 1000.0 / _abs_per_m)
# Rayleigh scattering ~ lambda^4, 70 m at 400 nm
water.set('scattering_length', 70e3 * (_wl / 400.0) ** 4)
water.density = 1.0

glass = Material('glass')
glass.set('refractive_index', 1.47 + 4.0e3 / _wl ** 2)
glass.set('absorption_length', 1000.0 * np.clip((_wl - 250.0) / 15.0, 0.1, 10.0))
glass.set('scattering_length', 1e6)
glass.density = 2.2

acrylic = Material('acrylic')
acrylic.set('refractive_index', 1.49)
acrylic.set('absorption_length', 2000.0)
acrylic.set('scattering_length', 1e6)

black_surface = Surface('black_surface')
black_surface.set('absorb', 1)

shiny_surface = Surface('shiny_surface')
shiny_surface.set('reflect_specular', 1)

lambertian_surface = Surface('lambertian_surface')
lambertian_surface.set('reflect_diffuse', 1)

# bialkali-like photocathode: QE peaks near 390 nm; the rest is shared between
# absorption and diffuse reflection so the three always sum to one
photocathode = Surface('photocathode')
_qe = 0.32 * np.exp(-0.5 * ((_wl - 390.0) / 75.0) ** 2)
photocathode.set('detect', _qe)
photocathode.set('absorb', _qe)
photocathode.set('reflect_diffuse', 1.0 - 2.0 * _qe)
