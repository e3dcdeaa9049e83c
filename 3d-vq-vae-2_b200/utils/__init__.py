"""Small stand-ins for the parts of the reference's `utils` package that the three caller scripts touch (the data
pipeline itself -- NRRD/LMDB/DICOM readers -- is out of scope): volume sources for the scripts and the
two on-disk formats they write."""
from .volumes import NpyVolumeDataset, SyntheticVolumeDataset, open_dataset  # noqa: F401
from .nrrd_io import write_nrrd, read_nrrd  # noqa: F401
