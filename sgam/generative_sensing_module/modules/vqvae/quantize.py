from sgam_neurips22_b200.quantize import VectorQuantizer2  # noqa: F401
