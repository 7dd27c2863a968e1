"""Reference import path: `from model_barf import ...` (barf/model_barf.py); the classes live in
model_camera_calibration.py."""
from .model_camera_calibration import BarfModel, CameraCalibrationModel, MipBarf, MipNeRF  # noqa: F401
