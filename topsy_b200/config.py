"""Module constants (same names and values as the reference's src/topsy/config.py:1-44; tests read them directly)."""

DEFAULT_RESOLUTION = 1024
DEFAULT_COLORMAP = 'twilight_shifted'

DEFAULT_SCALE = 200.0  # viewport half-width in kpc when the loader gives no better idea

TARGET_FPS = 30                        # interactive frames subsample particles to hold this rate
INITIAL_PARTICLES_TO_RENDER = 1e5      # first-frame particle budget
STATUS_LINE_UPDATE_INTERVAL = 0.2
STATUS_LINE_UPDATE_INTERVAL_RAPID = 0.05

GLIDE_TIME = 0.3

COLORBAR_ASPECT_RATIO = 0.15
COLORMAP_NUM_SAMPLES = 1000

TEST_DATA_NUM_PARTICLES_DEFAULT = int(1e6)

MAX_PARTICLES_PER_BUFFER = 2 ** 27          # particles per physical (split) buffer
MAX_PARTICLES_PER_EXPORT_RENDERCALL = 2 ** 25   # particles per render call in EXPORT frames

DEFAULT_CELLS_NSIDE = 16                    # cells per side of the spatial layout (nside^3 cells)
CELL_LAYOUT_FRACTIONAL_PADDING = 1e-5       # padding of the cell cube beyond the particle extent

JUPYTER_UI_LAG = 0.05

PROJECTED_DENSITY_NAME = "Projected density"

MAX_SURFACE_SMOOTH_PIXELS = 100
