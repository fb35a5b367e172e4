from sgam_neurips22_b200.model import VQModel  # noqa: F401
