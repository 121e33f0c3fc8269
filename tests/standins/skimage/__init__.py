"""Stand-in for the two scikit-image calls of the reference's host code (see ../README.md): `skimage.io.imread` (cv2) and
`skimage.measure.marching_cubes` (the generated-table marching cubes of oracle/marching_cubes.py).  Not scikit-image."""
from . import io, measure  # noqa: F401
__version__ = "0.0-standin"
