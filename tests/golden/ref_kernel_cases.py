"""The cases of tests/golden/ref_kernel_histories.npz: inputs are rebuilt from these parameters by
tests/scenes.py (NumPy Generator streams), so the fixture only holds the reference kernels' OUTPUTS.
Shared by make_golden_gpu.py (writes the fixture on a GPU box) and tests/test_oracle_physics.py."""

CASES = {
    # name: scene builder in tests/scenes.py + args, photon source, RNG seed, propagate arguments
    'sphere': dict(scene=('sphere_scene', (16,)), n=4000, src=dict(seed=2, wavelength=400.0), rng_seed=5,
                   max_steps=100, use_weights=False, scatter_first=0),
    'tiny': dict(scene=('tiny_detector', ()), n=6000, src=dict(seed=3, wl_range=(300, 600)), rng_seed=42,
                 max_steps=100, use_weights=False, scatter_first=0, daq_seed=5),
    'scint': dict(scene=('scintillator_scene', (12,)), n=6000, src=dict(seed=11, wl_range=(250, 450)), rng_seed=7,
                  max_steps=200, use_weights=False, scatter_first=0),
    'weights': dict(scene=('tiny_detector', ()), n=3000, src=dict(seed=5, wl_range=(350, 500)), rng_seed=3,
                    max_steps=50, use_weights=True, scatter_first=1),
    'wires': dict(scene=('wireplane_scene', ()), n=4000, src=dict(seed=21, wl_range=(350, 550), pos=(3.0, -7.0, -80.0)),
                  rng_seed=5, max_steps=100, use_weights=False, scatter_first=0),
    'one_step': dict(scene=('tiny_detector', ()), n=4000, src=dict(seed=9, wl_range=(300, 600)), rng_seed=8,
                     max_steps=1, use_weights=False, scatter_first=0),
}


def build(name):
    """(geometry, photons) of a case, from tests/scenes.py."""
    import scenes
    c = CASES[name]
    geo = getattr(scenes, c['scene'][0])(*c['scene'][1])
    return geo, scenes.point_source(c['n'], **c['src'])
