"""topsy_b200 -- B200-native SPH projection path behind topsy's Python surface.

    import topsy_b200 as topsy
    vis = topsy.test(1_000_000, render_resolution=1024)      # seeded synthetic snapshot, offscreen canvas
    vis.scale = 20.0; vis.rotate(0.0, 0.4); vis.quantity_name = "test-quantity"
    rgba = vis.get_sph_presentation_image()                   # (1024, 1024, 4) uint8
    content = vis.get_sph_image()                             # (1024, 1024) float32

Entry points mirror src/topsy/__init__.py of the reference: ``test``, ``load`` (``test://N`` or, with pynbody installed,
a simulation file) and ``topsy`` (an in-memory pynbody snapshot).
"""
from __future__ import annotations

__version__ = "0.1.0"

from . import config  # noqa: F401


def test(nparticle=config.TEST_DATA_NUM_PARTICLES_DEFAULT, **kwargs):
    from . import loader, visualizer
    return visualizer.Visualizer(data_loader_class=loader.TestDataLoader, data_loader_args=(nparticle,),
                                 data_loader_kwargs={'with_cells': kwargs.pop('with_cells', False),
                                                     'periodic': kwargs.get('periodic_tiling', False)},
                                 **kwargs)


def load(filename: str, center: str = "none", particle: str = "gas", resolution: int = config.DEFAULT_RESOLUTION,
         tile: bool = False, sphere_radius=None, sphere_center=None, render_mode: str = None):
    """``test://<N>`` builds the synthetic snapshot with N particles; anything else is loaded through pynbody."""
    from . import loader, visualizer
    if "test://" in filename:
        try:
            n_part = int(float(filename[7:]))
        except ValueError:
            n_part = config.TEST_DATA_NUM_PARTICLES_DEFAULT
        loader_class, loader_args = loader.TestDataLoader, (n_part,)
    else:
        import pynbody
        loader_class = loader.PynbodyDataLoader
        loader_args = (filename, center, particle)
        if sphere_radius is not None:
            region = pynbody.filt.Sphere(sphere_radius, sphere_center) if sphere_center is not None \
                else pynbody.filt.Sphere(sphere_radius)
            loader_args += (region,)
    return visualizer.Visualizer(data_loader_class=loader_class, data_loader_args=loader_args, periodic_tiling=tile,
                                 render_resolution=resolution, render_mode=render_mode or 'univariate')


def topsy(snapshot, quantity=None, **kwargs):
    """Visualizer for an in-memory pynbody snapshot (needs pynbody)."""
    from . import loader, visualizer
    vis = visualizer.Visualizer(data_loader_class=loader.PynbodyDataInMemory, data_loader_args=(snapshot,), **kwargs)
    vis.quantity_name = quantity
    return vis
